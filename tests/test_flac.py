"""libwb_audio.so: the FLAC decoder (include/wb_audio.h, csrc/audio/flac_decode.c) against an independent encoder
(tests/flac_writer.py) over every construct of the format, its checksum discipline, and the readers built on it
(`load_audio('.flac')`, `read_hf_dataset`).  Integer work: decoded PCM must be bit-exact.  No GPU."""
import ctypes
import hashlib
import os
import re

import numpy as np
import pytest

import flac_writer as FW
from conftest import ROOT
from whisper_trtllm_b200 import audio


def _signal(n, C=1, bps=16, seed=0, kind="speechlike"):
    rng = np.random.RandomState(seed)
    full = 1 << (bps - 1)
    if kind == "noise":
        return rng.randint(-full, full, size=(n, C)).astype(np.int64)
    t = np.arange(n)[:, None] / 16000.0
    x = 0.4 * np.sin(2 * np.pi * (180.0 + 40 * np.arange(C)[None, :]) * t) + 0.2 * np.sin(2 * np.pi * 1333.0 * t + np.arange(C)[None, :])
    x = x * (0.3 + 0.7 * np.abs(np.sin(3.0 * t))) + 0.01 * rng.randn(n, C)
    return np.clip(np.rint(x * full), -full, full - 1).astype(np.int64)


def _roundtrip(pcm, **kw):
    data = FW.encode(pcm, **kw)
    out, info = audio.decode_flac_pcm(data)
    want = np.asarray(pcm, dtype=np.int64).reshape(len(pcm), -1)
    assert out.dtype == np.int32 and out.shape == want.shape
    assert np.array_equal(out.astype(np.int64), want), kw
    return data, info


def test_library_exports_every_declared_symbol():
    src = open(os.path.join(ROOT, "include", "wb_audio.h")).read()
    declared = sorted(set(re.findall(r"WB_AUDIO_API\s+[\w\s\*]+?\b(wb_\w+)\s*\(", src)))
    assert declared == sorted(audio.AUDIO_SIGNATURES) and len(declared) == 5
    lib = audio.load_audio_lib()
    for s in declared:
        assert hasattr(lib, s)
    assert lib.wb_audio_version() >= 100
    assert ctypes.sizeof(audio.wb_flac_info) == 56        # the struct of the header, field for field


def test_md5_known_answers():
    lib = audio.load_audio_lib()
    for msg in (b"", b"a", b"abc", b"message digest", b"x" * 55, b"x" * 56, b"x" * 63, b"x" * 64, b"x" * 65, bytes(range(256)) * 40):
        out = ctypes.create_string_buffer(16)
        assert lib.wb_md5(msg, len(msg), out) == 0
        assert out.raw == hashlib.md5(msg).digest(), len(msg)
    assert hashlib.md5(b"abc").hexdigest() == "900150983cd24fb0d6963f7d28e17f72"      # RFC 1321 test suite


@pytest.mark.parametrize("kind", ["verbatim", ("fixed", 0), ("fixed", 1), ("fixed", 2), ("fixed", 3), ("fixed", 4),
                                  ("lpc", [1], 2, 0), ("lpc", [1800, -900], 12, 10), ("lpc", [-3, 7, 15, -16, 2, 1, 0, 9], 5, 3),
                                  ("lpc", list(range(-16, 16)), 15, 14)])
def test_subframe_types_mono_16bit(kind):
    pcm = _signal(5000, 1, 16, seed=1)
    data, info = _roundtrip(pcm, kind=kind, blocksize=1024)                  # 4 full blocks + a short last one
    assert info == {"sample_rate": 16000, "channels": 1, "bits_per_sample": 16, "total_samples": 5000, "min_blocksize": 1024,
                    "max_blocksize": 1024, "md5": hashlib.md5(pcm.astype("<i2").tobytes()).hexdigest()}


def test_constant_subframes_and_silence():
    _roundtrip(np.zeros((3000, 1), dtype=np.int64), kind="constant", blocksize=1024)
    _roundtrip(np.full((700, 2), -1234, dtype=np.int64), kind="constant", blocksize=256)
    # a long digital silence: far fewer than one bit per sample (the length is still found when STREAMINFO omits it)
    data = FW.encode(np.zeros((40000, 1), dtype=np.int64), kind="constant", blocksize=4096, total_samples_known=False, md5=False)
    assert len(data) < 200
    out, info = audio.decode_flac_pcm(data)
    assert info["total_samples"] == 0 and out.shape == (40000, 1) and not out.any()


@pytest.mark.parametrize("stereo", ["independent", "left_side", "side_right", "mid_side"])
@pytest.mark.parametrize("bps", [8, 12, 16, 20, 24])
def test_stereo_decorrelation_and_sample_sizes(stereo, bps):
    pcm = _signal(1500, 2, bps, seed=bps)
    pcm[0] = [-(1 << (bps - 1)), (1 << (bps - 1)) - 1]      # extreme values: the side channel needs bps + 1 bits
    pcm[1] = [(1 << (bps - 1)) - 1, -(1 << (bps - 1))]
    _roundtrip(pcm, bps=bps, stereo=stereo, kind=[("fixed", 1), "verbatim"], blocksize=512)
    _roundtrip(pcm, bps=bps, stereo=stereo, kind=("fixed", 2), blocksize=576, bits_from_streaminfo=True)


def test_32_bit_samples_and_side_channel_of_33_bits():
    pcm = _signal(600, 2, 32, seed=3, kind="noise")
    pcm[0] = [-(1 << 31), (1 << 31) - 1]
    for stereo in ("independent", "left_side", "side_right", "mid_side"):
        _roundtrip(pcm, bps=32, stereo=stereo, kind="verbatim", blocksize=256)
    _roundtrip(_signal(600, 1, 32, seed=4), bps=32, kind=("fixed", 2), blocksize=256, five_bit=True)


@pytest.mark.parametrize("C", [1, 2, 3, 6, 8])
def test_channel_counts(C):
    pcm = _signal(900, C, 16, seed=C)
    _roundtrip(pcm, kind=("fixed", 3), blocksize=192)


@pytest.mark.parametrize("partition_order,five_bit,escape", [(0, False, False), (3, False, False), (6, False, False), (2, True, False),
                                                               (0, False, True), (4, False, True), (3, True, True)])
def test_residual_partitions_rice_parameters_and_escapes(partition_order, five_bit, escape):
    pcm = _signal(4096 + 1024, 1, 16, seed=5)
    pcm[100:164] = 0                                      # an all-zero partition: escape code with 0 bits per sample
    _roundtrip(pcm, kind=("fixed", 2), blocksize=1024, partition_order=partition_order, five_bit=five_bit, escape=escape)
    _roundtrip(pcm, kind=("lpc", [3, -1], 4, 1), blocksize=1024, partition_order=partition_order, five_bit=five_bit, escape=escape)


def test_large_residuals_need_five_bit_parameters():
    pcm = _signal(2048, 1, 24, seed=6, kind="noise")      # white noise at 24 bits: Rice parameters above 14
    _roundtrip(pcm, bps=24, kind=("fixed", 0), blocksize=1024, five_bit=True)
    _roundtrip(pcm, bps=24, kind=("fixed", 4), blocksize=1024, five_bit=True, partition_order=2)


def test_wasted_bits():
    pcm = _signal(2000, 2, 16, seed=7) * 8                # three zero low bits in every sample
    pcm = np.clip(pcm, -32768, 32760)
    pcm -= pcm % 8
    for kind in ("verbatim", ("fixed", 2), ("lpc", [2, -1], 3, 0)):
        _roundtrip(pcm, kind=kind, blocksize=500, wasted="auto", stereo="mid_side")
        _roundtrip(pcm, kind=kind, blocksize=500, wasted=2)
    _roundtrip(np.full((300, 1), 4096, dtype=np.int64), kind="constant", blocksize=300, wasted="auto")


@pytest.mark.parametrize("blocksize", [16, 192, 255, 256, 257, 576, 1000, 1152, 2304, 4096, 4608, 8192, 16384, 32768, 65535])
def test_block_size_codes(blocksize):
    pcm = _signal(blocksize + 37, 1, 16, seed=8)
    _roundtrip(pcm, kind=("fixed", 1), blocksize=blocksize)
    _roundtrip(pcm, kind=("fixed", 1), blocksize=blocksize, explicit_blocksize=True)


@pytest.mark.parametrize("rate,mode", [(16000, "code"), (16000, "streaminfo"), (16000, "khz"), (16000, "hz"), (44100, "code"),
                                       (96000, "code"), (11025, "hz"), (250000, "khz"), (655350, "tens"), (88200, "code"),
                                       (8000, "code"), (22050, "tens")])
def test_sample_rate_codes(rate, mode):
    data, info = _roundtrip(_signal(400, 1, 16, seed=9), rate=rate, rate_mode=mode, blocksize=256)
    assert info["sample_rate"] == rate
    x, r = audio.decode_flac(data)
    assert r == rate and x.dtype == np.float32


def test_variable_block_size_stream_and_long_frame_numbers():
    pcm = _signal(6000, 1, 16, seed=10)
    per_frame = {0: dict(blocksize=1000), 1: dict(blocksize=24, kind="verbatim"), 2: dict(blocksize=4096, kind=("lpc", [2, -1], 3, 0)),
                 3: dict(blocksize=16, partition_order=2, kind=("fixed", 0))}
    _roundtrip(pcm, variable=True, blocksize=4096, per_frame=per_frame)
    # coded frame numbers of 2..7 bytes
    for first in (0x7f, 0x80, 0x7ff, 0x800, 0xffff, 0x10000, 0x1fffff, 0x200000, 0x3ffffff, 0x4000000, 0x7ffffff0):
        _roundtrip(pcm[:700], blocksize=256, first_frame_number=first)
    big = FW.utf8_number((1 << 36) - 1)
    assert len(big) == 7 and big[0] == 0xfe


def test_metadata_layouts_and_trailing_bytes():
    pcm = _signal(1000, 1, 16, seed=11)
    _roundtrip(pcm, blocksize=256, padding_block=False)
    _roundtrip(pcm, blocksize=256, id3=True)
    _roundtrip(pcm, blocksize=256, trailer=b"TAG" + b"\x00" * 125)                  # ID3v1 tag after the last frame
    _roundtrip(pcm, blocksize=256, md5=False)                                        # no signature: nothing to verify
    data = FW.encode(pcm, blocksize=256, total_samples_known=False)
    out, info = audio.decode_flac_pcm(data)
    assert info["total_samples"] == 0 and np.array_equal(out[:, 0], pcm[:, 0])


def test_float_waveform_matches_soundfile_convention(tmp_path):
    pcm = _signal(3000, 2, 16, seed=12)
    data = FW.encode(pcm, stereo="mid_side", blocksize=1024)
    x, rate = audio.decode_flac(data)
    assert rate == 16000 and x.dtype == np.float32
    assert np.array_equal(x, (pcm.mean(axis=1) / 32768.0).astype(np.float32))       # int / 2^(bits-1), channels averaged
    p = tmp_path / "utt.flac"
    p.write_bytes(FW.encode(pcm[:, :1], blocksize=1024))
    assert np.array_equal(audio.load_audio(str(p)), (pcm[:, 0] / 32768.0).astype(np.float32))
    (tmp_path / "utt8k.flac").write_bytes(FW.encode(pcm[:, :1], rate=8000))
    with pytest.raises(ValueError, match="8000"):
        audio.load_audio(str(tmp_path / "utt8k.flac"))


# ---------------------------------------------------------------------------------------------------- corruption is detected
def _decode_error(data, verify_md5=True):
    with pytest.raises(audio.AudioDecodeError) as e:
        audio.decode_flac_pcm(bytes(data), verify_md5)
    return e.value


def test_every_single_bit_flip_in_the_audio_is_detected():
    pcm = _signal(600, 2, 16, seed=13)
    data = FW.encode(pcm, stereo="left_side", kind=("fixed", 2), blocksize=256, partition_order=1)
    start = len(FW.encode(pcm[:0], blocksize=256))          # metadata only: the first frame starts here
    rng = np.random.RandomState(0)
    for pos in rng.choice(np.arange(start, len(data)), size=150, replace=False):
        bad = bytearray(data)
        bad[pos] ^= 1 << rng.randint(8)
        err = _decode_error(bad)
        assert err.code in (-2, -3), (pos, err)


def test_checksum_and_format_errors_are_loud():
    pcm = _signal(600, 1, 16, seed=14)
    data = bytearray(FW.encode(pcm, blocksize=256))
    assert _decode_error(b"RIFF" + bytes(40)).code == -2
    assert _decode_error(data[:60]).code == -2                                       # cut inside the first frame
    assert _decode_error(data[:-1]).code in (-2, -3)
    assert _decode_error(data[:len(data) // 2]).code in (-2, -3)
    # wrong MD5 in STREAMINFO: frames are fine, the signature check fails — and only when it is asked for
    bad = bytearray(data)
    bad[4 + 4 + 18] ^= 0xff
    assert _decode_error(bad).code == -3
    out, _ = audio.decode_flac_pcm(bytes(bad), verify_md5=False)
    assert np.array_equal(out[:, 0], pcm[:, 0])
    # a PCM change that keeps every frame CRC valid is still caught by the MD5: re-encode different audio under the old signature
    other = bytearray(FW.encode(pcm + (np.arange(600)[:, None] == 300), blocksize=256))
    other[8 + 18:8 + 34] = data[8 + 18:8 + 34]
    assert "MD5" in str(_decode_error(other))
    # buffer too small
    n = ctypes.c_uint64()
    buf = np.empty((100, 1), dtype=np.int32)
    lib = audio.load_audio_lib()
    assert lib.wb_flac_decode_i32(bytes(data), len(data), buf.ctypes.data_as(ctypes.c_void_p), 100, ctypes.byref(n), 1) == -1
    assert b"too small" in lib.wb_audio_last_error() and n.value == 0
    assert lib.wb_flac_read_info(None, 0, None) == -1


def test_reserved_fields_are_rejected():
    pcm = _signal(300, 1, 16, seed=15)
    good = FW.encode(pcm, blocksize=256, padding_block=False)
    first = 4 + 4 + 34

    def patched(offset, value, fix_crc8=True):
        b = bytearray(good)
        b[first + offset] = value
        if fix_crc8:      # header: sync(2) codes(2) number(1) crc8(1) for frame 0 with standard codes
            b[first + 5] = FW.crc(bytes(b[first:first + 5]), 0x07, 8)
        return b
    assert "sync" in str(_decode_error(patched(1, 0xfa, False)))           # reserved bit after the sync code
    assert _decode_error(patched(2, 0x05)).code == -2           # block size code 0 (reserved)
    assert _decode_error(patched(2, 0xcf)).code == -2           # sample rate code 15 (invalid)
    assert _decode_error(patched(3, 0xb8)).code == -2           # channel assignment 11 (reserved)
    assert _decode_error(patched(3, 0x06)).code == -2           # sample size code 3 (reserved)
    assert _decode_error(patched(3, 0x09)).code == -2           # reserved bit
    assert _decode_error(patched(3, 0x18)).code == -4           # stereo frame in a mono stream: unsupported, said so


# ---------------------------------------------------------------------------------------------------- HF datasets directory
def test_hf_dataset_directory_with_flac_bytes(tmp_path):
    datasets = pytest.importorskip("datasets")
    waves = [_signal(4000 + 500 * i, 1, 16, seed=20 + i) for i in range(3)]
    rows = [{"bytes": FW.encode(w, blocksize=4096, kind=("lpc", [2, -1], 3, 0), partition_order=3), "path": f"1272-128104-{i:04d}.flac"}
            for i, w in enumerate(waves)]
    texts = ["MISTER QUILTER IS THE APOSTLE", "NOR IS MISTER QUILTER'S MANNER LESS INTERESTING", "HE TELLS US"]
    ds = datasets.Dataset.from_dict({"audio": rows, "text": texts, "id": ["a", "b", "c"]})
    ds = ds.cast_column("audio", datasets.Audio(sampling_rate=16000, decode=False))
    ds.save_to_disk(str(tmp_path / "librispeech_asr_dummy"))
    assert audio.is_hf_dataset(str(tmp_path / "librispeech_asr_dummy")) and not audio.is_hf_dataset(str(tmp_path))
    got, refs = audio.read_hf_dataset(str(tmp_path / "librispeech_asr_dummy"))
    assert refs == texts and len(got) == 3
    for g, w in zip(got, waves):
        assert g.dtype == np.float32 and np.array_equal(g, (w[:, 0] / 32768.0).astype(np.float32))
    with pytest.raises(KeyError):
        audio.read_hf_dataset(str(tmp_path / "librispeech_asr_dummy"), audio_column="speech")


# ---------------------------------------------------------------------------------------------------- memory safety
def test_mutation_fuzz_under_address_and_ub_sanitizers(tmp_path):
    """tests/fuzz_flac.c + the decoder compiled with -fsanitize=address,undefined: 15 000 mutated streams (bit flips, random
    bytes, truncations, 0x00 / 0xff runs); any out-of-bounds access or undefined behaviour aborts the run."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "fuzz_flac")
    build = subprocess.run([gcc, "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-DWB_AUDIO_BUILD", "-o", exe,
                            os.path.join(ROOT, "tests", "fuzz_flac.c"),
                            os.path.join(ROOT, "whisper_trtllm_b200", "csrc", "audio", "flac_decode.c")], capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr.lower() + build.stdout.lower():
        pytest.skip("sanitizer runtime not available")
    assert build.returncode == 0, build.stderr[-2000:]
    cases = [
        (_signal(3000, 2, 16, 1), dict(stereo="mid_side", kind=("lpc", [1800, -900], 12, 10), blocksize=1024, partition_order=3)),
        (_signal(2000, 1, 24, 2), dict(bps=24, kind=("fixed", 4), blocksize=576, five_bit=True, partition_order=2)),
        (_signal(1500, 3, 8, 3), dict(bps=8, kind="verbatim", blocksize=255)),
        (np.clip(_signal(2500, 2, 16, 4) * 4, -32768, 32764), dict(stereo="left_side", kind=("fixed", 2), blocksize=500, escape=True,
                                                                     partition_order=2, variable=True)),
        (np.zeros((5000, 1), dtype=np.int64), dict(kind="constant", blocksize=4096, total_samples_known=False, md5=False)),
    ]
    files = []
    for i, (pcm, kw) in enumerate(cases):
        files.append(str(tmp_path / f"case{i}.flac"))
        with open(files[-1], "wb") as f:
            f.write(FW.encode(pcm, **kw))
    run = subprocess.run([exe, "3000"] + files, capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, (run.stdout + run.stderr)[-3000:]
    counts = dict(zip(run.stdout.split()[0::2], map(int, run.stdout.split()[1::2])))
    assert sum(counts.values()) == 15000 and counts["format"] + counts["checksum"] > 10000


# ---------------------------------------------------------------------------------------------------- known-answer files
def test_rfc9639_example_files():
    """The three example files of RFC 9639 Appendix D (made by reference libFLAC 1.3.3): the decoder reproduces their PCM, and the
    files certify themselves — hashlib's MD5 of the decoded samples equals the signature stored inside each file."""
    import json
    with open(os.path.join(ROOT, "tests", "golden", "flac_rfc9639.json")) as f:
        g = json.load(f)
    assert len(g["files"]) == 3
    for case in g["files"]:
        data = bytes.fromhex(case["hex"])
        pcm, info = audio.decode_flac_pcm(data, verify_md5=True)
        assert (info["sample_rate"], info["channels"], info["bits_per_sample"]) == (case["sample_rate"], case["channels"], case["bits_per_sample"])
        assert pcm.tolist() == case["pcm"], case["name"]
        nbytes = (info["bits_per_sample"] + 7) // 8
        raw = b"".join(int(v).to_bytes(nbytes, "little", signed=True) for v in pcm.reshape(-1))
        assert hashlib.md5(raw).hexdigest() == info["md5"] == case["md5"]
    first = g["files"][0]
    assert first["pcm"] == [[25588, 10416]] and first["sample_rate"] == 44100

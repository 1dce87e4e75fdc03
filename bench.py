#!/usr/bin/env python
"""bench.py — RTFx (audio-seconds transcribed per second) of whisper-medium.en greedy decoding on B200.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.
A "step" = one pass of the hot path over one batch of synthetic input: encoder + cross-K/V projection +
the full 447-step greedy loop.  Default = BASELINE.json configs[3]: a GLOBAL batch of 256 30-second utterances
sharded data-parallel over the N ranks (256 / N per rank, "scaling": "strong"; utterances are independent, no
data-path collective; the only collective is the final NCCL all-gather of the token ids).  `--batch B` switches
to B utterances PER GPU ("scaling": "weak"); under N > 1 the default line also carries that weak-scaling figure
(256 per GPU, a short extra measurement) as `weak_scaling`.

  value   whole-job throughput with the log-mel inputs already resident in HBM (CUDA events, max over ranks)
  e2e     the same metric through the public host-buffer call (pinned host log-mel -> H2D -> encode ->
          greedy -> D2H of the ids) inside the timed region
  roofline   the dominant kernel (decode cross-attention, HBM-bound) timed live with CUDA events around
             every one of its launches inside the timed region
  cpu_baseline   the reference's CPU path (oracle port, fp32 torch) on this box's host cores, bounded sample

`--impl reference` times only the CPU path (rank 0) with the same metric/config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "RTFx (audio-s/s) medium.en greedy"
UNIT = "audio-s/s"
AUDIO_SECONDS = 30.0


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--size", default="medium.en")
    p.add_argument("--batch", type=int, default=0, help="utterances PER GPU (weak scaling); default 0 = shard --global-batch over the ranks")
    p.add_argument("--global-batch", type=int, default=256, help="utterances in total, sharded over the ranks (strong scaling, BASELINE configs[3])")
    p.add_argument("--no-weak", action="store_true", help="skip the extra weak-scaling measurement of a multi-GPU strong-scaling run")
    p.add_argument("--no-probe", action="store_true", help="skip the parity probe against the CPU port")
    p.add_argument("--no-microbench", action="store_true", help="skip the decode-step microbench (BASELINE configs[4])")
    p.add_argument("--chain", type=int, default=-1, help="(dev) wb_set_decode_chain_path")
    p.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    p.add_argument("--enc-chunk", type=int, default=32)
    p.add_argument("--max-length", type=int, default=448)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--pdl", action="store_true", help="(dev) enable programmatic dependent launch between decode kernels")
    p.add_argument("--no-graph", action="store_true", help="(dev) launch every decode kernel individually instead of replaying a CUDA graph")
    p.add_argument("--streams", type=int, default=1, help="sub-batches decoded concurrently on separate streams (per GPU)")
    p.add_argument("--bulk-attn", action="store_true", help="(dev) cp.async.bulk ring kernel for the cross attention")
    p.add_argument("--self-attn-variant", type=int, default=0, help="(dev) paged self-attention kernel variant 0..4")
    p.add_argument("--attn-variant", type=int, default=0, help="(dev) cross-attention kernel variant 2..6 (threads, unroll)")
    p.add_argument("--breakdown-only", default="", help="(dev) comma-separated kernel classes for --breakdown (default: all)")
    p.add_argument("--breakdown", action="store_true", help="(dev) per-kernel-class device time of one extra step, to stderr")
    return p.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


_CPU_WEIGHTS = {}
# (utterances, decode steps) of one CPU sample with its rough cost in seconds on 16 host cores (encoder ~1 s per utterance,
# decode step ~25 ms + 2.5 ms per utterance): the largest one that keeps K + W passes within ~5 minutes is used
CPU_SAMPLES = [((16, 128), 26.0), ((8, 128), 14.0), ((8, 64), 11.0), ((4, 128), 8.5), ((4, 64), 6.5)]


def pick_cpu_sample(passes, budget_s=300.0):
    per_pass = budget_s / max(passes, 1)
    for shape, cost in CPU_SAMPLES:
        if cost <= per_pass:
            return shape
    return CPU_SAMPLES[-1][0]


def cpu_reference_rtfx(size, max_length, sample_batch=16, sample_steps=128):
    """The reference's CPU path (oracle port, fp32) on this host: encoder + `sample_steps` greedy decode steps for
    `sample_batch` utterances are RUN and timed; the remaining max_length-1-sample_steps steps are extrapolated at the mean
    step time of the second half of the sample (later steps attend over longer caches, so this does not flatter the CPU).
    -> dict(value, cores, sample, measured_s, extrapolated_s)."""
    from oracle import synth, whisper_ref as R
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = synth.make_config(size, max_length=max_length)
    key = (size, max_length)
    if _CPU_WEIGHTS.get("key") != key:        # synthetic weights are built once per process, outside the timed parts
        _CPU_WEIGHTS.update(key=key, sd=synth.make_weights(cfg, seed=0))
    sd = _CPU_WEIGHTS["sd"]
    mel = synth.make_mel(sample_batch, seed=1234)
    sample_steps = min(sample_steps, max_length - 1)
    with torch.no_grad():
        t0 = time.perf_counter()
        enc = R.encode(mel, sd, cfg)
        t_enc = time.perf_counter() - t0
        ids = torch.full((sample_batch, 1), cfg["decoder_start_token_id"], dtype=torch.long)
        past = None
        stamps = [time.perf_counter()]
        for n in range(sample_steps):
            logits, past = R.decoder_forward(ids[:, -1:], enc, sd, cfg, past)
            nxt = R.process_logits(logits[:, -1, :], ids.shape[1], cfg).argmax(-1)
            ids = torch.cat([ids, nxt[:, None]], dim=-1)
            stamps.append(time.perf_counter())
    t_dec = stamps[-1] - stamps[0]
    half = sample_steps // 2
    late_step = (stamps[-1] - stamps[half]) / max(sample_steps - half, 1)
    measured = t_enc + t_dec
    total = measured + late_step * (max_length - 1 - sample_steps)
    value = AUDIO_SECONDS * sample_batch / total
    sample = (f"{size} fp32 oracle port, batch {sample_batch}: encoder {t_enc:.2f}s + {sample_steps} greedy decode steps "
              f"{t_dec:.2f}s RUN ({measured:.2f}s measured); the other {max_length - 1 - sample_steps} steps extrapolated at "
              f"{late_step * 1e3:.1f} ms/step (mean of the sample's second half) -> {total:.2f}s per {sample_batch} utterances")
    return {"value": value, "cores": cores, "sample": sample, "measured_s": measured, "extrapolated_s": total,
            "sample_batch": sample_batch, "sample_steps": sample_steps}


def job_ceiling(cfg, batch, elem_bytes, peaks):
    """Whole-job roofline of one step: encoder + cross-K/V projection at the measured sustained tensor
    throughput, then max_length-1 decode steps at the measured HBM bandwidth, each streaming the decoder weights once for the
    batch and, per utterance, the cross K/V of every layer plus the self K/V written so far.  -> dict with the ceiling RTFx."""
    d, S, V = cfg["d_model"], cfg["max_source_positions"], cfg["vocab_size"]
    Le, Ld, steps = cfg["encoder_layers"], cfg["decoder_layers"], cfg["max_length"] - 1
    mel, frames = cfg["num_mel_bins"], 2 * S
    enc_flops = 2 * frames * 3 * mel * d + 2 * S * 3 * d * d + Le * (24 * S * d * d + 4 * S * S * d)
    xkv_flops = Ld * 4 * S * d * d
    w_step = Ld * 14 * d * d + V * d                      # decoder weights read per step (cross k/v projections excluded)
    cross = 2 * Ld * S * d                                # cross K/V elements per utterance, read every step
    self_kv = Ld * d * steps * (steps + 1)                # sum over t = 1..steps of 2 * Ld * t * d
    dec_bytes = elem_bytes * (steps * w_step + batch * (steps * cross + self_kv))
    tensor_tflops = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops") or 1400.0   # inside a long step: sustained
    t_enc = batch * (enc_flops + xkv_flops) / (tensor_tflops * 1e12)
    t_dec = dec_bytes / (peaks["hbm_gbs"] * 1e9)
    return {"ceiling": round(AUDIO_SECONDS * batch / (t_enc + t_dec), 1), "unit": UNIT, "encoder_floor_ms": round(t_enc * 1e3, 1),
            "decode_floor_ms": round(t_dec * 1e3, 1), "flops_per_utterance": enc_flops + xkv_flops,
            "decode_bytes_per_step_mean": int(dec_bytes / steps),
            "peaks": {"tensor_tflops": tensor_tflops, "hbm_gbs": peaks["hbm_gbs"]}}


def batch_plan(args, world, rank):
    """-> (rows of this rank, largest shard, global batch, scaling)."""
    from whisper_trtllm_b200 import dp
    if args.batch > 0:
        return args.batch, args.batch, args.batch * world, "weak"
    b, e = dp.shard_range(args.global_batch, world, rank)
    return e - b, dp.max_shard(args.global_batch, world), args.global_batch, "strong"


def workload_config(args, world):
    """The `config` object of the JSON line (same for our arm and the reference arm)."""
    _, per_gpu, total, scaling = batch_plan(args, world, 0)
    how = (f"global batch {total} x 30 s synthetic log-mel sharded over {world} GPU(s) ({per_gpu} per GPU)" if scaling == "strong"
           else f"batch {per_gpu} x 30 s synthetic log-mel per GPU")
    return {"workload": f"whisper-{args.size} {args.dtype} greedy, {how}, {args.max_length}-token max decode, "
                        "data-parallel by utterance (BASELINE.json configs[3])",
            "size": args.size, "batch_per_gpu": per_gpu, "global_batch": total, "max_length": args.max_length,
            "l2": f"inputs_exceed_l2 (per step the kernels stream ~{0.19 * per_gpu:.0f} GB of KV cache + 0.8 GB of weights)",
            "parallelism": f"dp{world}"}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; /root/reference is Python + closed TensorRT and cannot be
    installed) on rank 0 with every host thread.  One "step" = one bounded sample of the workload: the sample
    is RUN (encoder + greedy steps), `ms_per_step` is its measured wall time; `value` extrapolates the rest of the 447-step
    loop (stated in cpu_baseline.sample)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    sb, ss = pick_cpu_sample(args.warmup + args.steps)
    runs = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_rtfx(args.size, args.max_length, sb, ss)
        log(f"[reference] pass {i}: {r['value']:.2f} {UNIT} ({r['measured_s']:.2f}s measured)")
        if i >= args.warmup:
            runs.append(r)
    value = sum(r["value"] for r in runs) / len(runs)
    ms = 1e3 * sum(r["measured_s"] for r in runs) / len(runs)
    _, _, _, scaling = batch_plan(args, world, 0)
    last = runs[-1]
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 1), "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": last["cores"], "kind": "port", "sample": last["sample"],
                         "measured_ms_per_sample": round(ms, 1),
                         "extrapolated_ms_per_sample": round(1e3 * sum(r["extrapolated_s"] for r in runs) / len(runs), 1)},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "ms_per_step is the MEASURED wall time of one sample (what actually ran); value = 30 s x sample utterances / "
                "(measured + extrapolated remainder of the 447-step loop).  One CPU process regardless of n_gpus.",
    }
    emit(line)


def timed_passes(eng, mel_dev, mel_host, ids_host, steps, gather, sync_all, prof_step):
    """The two timed regions of one engine: (device ms with resident inputs, wall ms with host buffers, launches, cross-attention
    event ms, launches timed)."""
    def step_device():
        ids = eng.generate(mel_dev)
        gather(ids)
        return ids

    def step_e2e():
        mel = mel_host.to(mel_dev.device, non_blocking=True)
        ids = eng.generate(mel)
        gather(ids)
        ids_host[:ids.shape[0], :ids.shape[1]].copy_(ids)  # D2H of the result (blocking)
        return ids

    if prof_step >= 0:
        eng.profile("cross_attn", decode_step=prof_step)
    launches0 = eng.launch_count()
    sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        ids = step_device()
    ev1.record()
    sync_all()
    dev_ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0
    xattn_ms, xattn_n = (eng.profile_read() if prof_step >= 0 else (0.0, 0))
    eng.profile(None)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_e2e()
    sync_all()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    return dev_ms, e2e_ms, launches, xattn_ms, xattn_n, ids


def parity_probe(eng, cfg, mel_host, mel_dev, free_ids, size, max_length, rows=4, steps=8):
    """Teacher-force the BENCHMARKED engine (full batch, same kernels) on `rows` of its rows for `steps` steps with the ids of
    the CPU port (oracle/whisper_ref.py = the reference's algorithm, pinned to the real reference by tests/golden) and
    compare the logits of those rows: max error relative to max |logit|, argmax agreement on the free steps."""
    from oracle import synth, whisper_ref as R
    B = mel_dev.shape[0]
    rows = min(rows, B)
    steps = min(steps, max_length - 1)
    key = (size, max_length)
    if _CPU_WEIGHTS.get("key") != key:
        _CPU_WEIGHTS.update(key=key, sd=synth.make_weights(cfg, seed=0))
    sd = _CPU_WEIGHTS["sd"]
    cfg_s = dict(cfg, max_length=steps + 1)
    torch.set_num_threads(os.cpu_count() or 1)
    ref_ids, _, ref_logits = R.greedy(mel_host[:rows].clone(), sd, cfg_s, return_logits=True)
    forced = free_ids[:, :steps + 1].clone().long().cpu()
    if forced.shape[1] < steps + 1:
        return {"skipped": "the loop stopped before the probe length"}
    forced[:rows] = ref_ids
    ids, logits = eng.generate(mel_dev, max_new_tokens=steps, forced_tokens=forced, dump_logits_steps=steps)
    max_rel, agree, total = 0.0, 0, 0
    for s in range(steps):
        got = logits[s][:rows].float().cpu()
        want = ref_logits[s]
        max_rel = max(max_rel, float((got - want).abs().max() / want.abs().max()))
        if s >= 1:   # step 0 is the forced token
            agree += int((R.process_logits(got, s + 1, cfg).argmax(-1) == ref_ids[:, s + 1]).sum())
            total += rows
    return {"rows": rows, "steps": steps, "batch": B, "max_rel": round(max_rel, 5), "argmax_agree": round(agree / max(total, 1), 4),
            "tolerance": 3e-2, "against": "fp32 CPU port of the reference (oracle/whisper_ref.py), teacher-forced"}


def decode_step_microbench(eng, cfg, B_max, es, peaks, t_len=224, steps=8):
    """BASELINE.json configs[4]: device time of ONE greedy decode step (us) at sequence length ~t_len for batch 1..B_max on the
    benchmarked engine (cross K/V of every row already projected), CUDA-graph replay, CUDA events around `steps` steps."""
    d, L = cfg["d_model"], cfg["decoder_layers"]
    w_step = L * 14 * d * d + cfg["vocab_size"] * d
    out = {}
    for B in [b for b in (1, 8, 16, 32, 64, 128, 256, 512) if b <= B_max]:
        eng.decode_begin(B)
        eng.decode_run(max_steps=t_len - 1, check_every=1 << 20)
        eng.decode_run(max_steps=2, check_every=1 << 20)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        eng.decode_run(max_steps=steps, check_every=1 << 20)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / steps * 1e3
        tt = t_len + 2 + steps / 2
        floor = es * (w_step + B * (2 * L * cfg["max_source_positions"] * d + 2 * L * tt * d)) / (peaks["hbm_gbs"] * 1e9) * 1e6
        out[str(B)] = {"us": round(us, 1), "hbm_floor_us": round(floor, 1), "frac": round(floor / us, 3)}
    return {"length": t_len, "unit": "us per decode step (24 layers + LM head + argmax)", "by_batch": out}


_REAL_STDOUT = None


def claim_stdout():
    """Rank 0 prints ONE JSON line on stdout.  Libraries write there too (NCCL's version banner goes to the C stdout): file
    descriptor 1 is pointed at stderr for the whole run and the line is written to the saved descriptor at the end."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def main():
    args = parse_args()
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    if args.warmup < 3 and not os.environ.get("WB_BENCH_DEV"):
        args.warmup = 3  # timing rules: at least 3 warm-up steps

    from whisper_trtllm_b200 import synthetic as synth  # seeded synthetic weights / inputs (bit-reproducible)
    from whisper_trtllm_b200 import WhisperEngine

    from whisper_trtllm_b200 import _abi
    if args.pdl:
        _abi.call("wb_set_pdl", 1)
    if args.no_graph:
        _abi.call("wb_set_cuda_graphs", 0)
    if args.bulk_attn:
        _abi.call("wb_set_decode_attention_backend", 1)
    if args.self_attn_variant:
        _abi.call("wb_set_self_attention_warp_kernel", args.self_attn_variant)
    if args.attn_variant:
        _abi.call("wb_set_decode_attention_backend", args.attn_variant)
    if args.chain >= 0:
        _abi.call("wb_set_decode_chain_path", args.chain)
    B, B_max, total, scaling = batch_plan(args, world, rank)
    assert B > 0, "every rank needs at least one utterance"
    cfg = synth.make_config(args.size, max_length=args.max_length)
    t0 = time.time()
    sd = synth.make_weights(cfg, seed=0)
    eng = WhisperEngine(cfg, sd, dtype=args.dtype, max_batch=B, enc_chunk=min(args.enc_chunk, B), device=dev, n_streams=args.streams)
    log(f"[rank {rank}] weights packed in {time.time() - t0:.1f}s; {B} utterances on this rank; workspace {eng.workspace.numel() / 2**30:.1f} GiB")
    # the global batch is ONE seeded set of utterances; a rank holds its contiguous shard (weak mode: a different seed per rank)
    if scaling == "strong":
        from whisper_trtllm_b200 import dp
        b0, b1 = dp.shard_range(total, world, rank)
        mel_host = synth.make_mel(total, seed=1234)[b0:b1].clone().pin_memory()
    else:
        mel_host = synth.make_mel(B, seed=1234 + rank).pin_memory()
    mel_dev = mel_host.to(dev)
    ids_host = torch.empty(B, cfg["max_target_positions"], dtype=torch.int32).pin_memory()
    from whisper_trtllm_b200 import dp

    def make_gather(n_total):
        def gather(ids):
            if world > 1:  # the path's only collective: final token gather over NCCL
                return dp.gather_tokens(ids, n_total, args.max_length, cfg["pad_token_id"])
            return ids
        return gather

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    gather = make_gather(total)
    for i in range(args.warmup):
        t0 = time.time()
        if i % 2 == 0:
            ids = eng.generate(mel_dev)
        else:
            ids = eng.generate(mel_host.to(dev, non_blocking=True))
            ids_host[:, :ids.shape[1]].copy_(ids)
        gather(ids)
        torch.cuda.synchronize()
        log(f"[rank {rank}] warmup {i}: {time.time() - t0:.2f}s, ids {tuple(ids.shape)}")

    sampler = ClockSampler(local_rank)
    # live roofline timing inside the timed region: CUDA events around every cross-attention launch of ONE decode step per
    # greedy loop (the middle one; it is launched eagerly, the other 446 steps replay the CUDA graph).  The kernel's work
    # does not depend on the step (always 1500 keys), so the sample is representative; ncu shares are under profiles/.
    prof_step = (args.max_length - 1) // 2
    sampler.start()
    dev_ms, e2e_ms, launches, xattn_ms, xattn_n, free_ids = timed_passes(eng, mel_dev, mel_host, ids_host, args.steps, gather,
                                                                         sync_all, prof_step)
    clocks = sampler.stop()

    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = t.tolist()
    audio_s = AUDIO_SECONDS * total * args.steps
    value = audio_s / (dev_ms / 1e3)
    e2e_value = audio_s / (e2e_ms / 1e3)

    # phases of one extra identical pass (not part of the timed regions): encoder + cross-K/V projection vs greedy loop
    pe = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    sync_all()
    pe[0].record()
    eng.encode(mel_dev, return_hidden=False)
    pe[1].record()
    eng.greedy(B)
    pe[2].record()
    torch.cuda.synchronize()
    enc_ms, dec_ms = pe[0].elapsed_time(pe[1]), pe[1].elapsed_time(pe[2])

    if args.breakdown and rank == 0:
        only = [c for c in args.breakdown_only.split(",") if c]
        for cls in (only or eng.PROF_CLASSES):
            eng.profile(cls)   # every step, eager launches
            eng.generate(mel_dev)
            ms, n = eng.profile_read()
            log(f"[breakdown] {cls:10s} {ms:9.1f} ms over {n} launches ({ms / max(n, 1) * 1e3:.1f} us each)")
        eng.profile(None)

    peaks, peak_src = measured_peaks()
    es = 2 if args.dtype == "bf16" else 4
    probe = micro = None
    if rank == 0 and args.streams == 1:
        if not args.no_probe:
            try:
                probe = parity_probe(eng, cfg, mel_host, mel_dev, free_ids, args.size, args.max_length)
                log(f"[rank 0] parity probe: {probe}")
            except Exception as e:   # the probe never hides a number: it reports its own failure
                probe = {"error": repr(e)}
        if not args.no_microbench and world == 1:
            micro = decode_step_microbench(eng, cfg, B, es, peaks)
            log(f"[rank 0] decode-step microbench: {micro}")

    # ---- weak-scaling figure of a multi-GPU strong-scaling run: 256 utterances PER GPU on a second engine, fewer passes
    weak = None
    if world > 1 and scaling == "strong" and not args.no_weak:
        eng.close()
        del eng
        torch.cuda.empty_cache()
        Bw = total
        engw = WhisperEngine(cfg, sd, dtype=args.dtype, max_batch=Bw, enc_chunk=min(args.enc_chunk, Bw), device=dev)
        melw_host = synth.make_mel(Bw, seed=1234 + rank).pin_memory()
        melw = melw_host.to(dev)
        idsw_host = torch.empty(Bw, cfg["max_target_positions"], dtype=torch.int32).pin_memory()
        gw = make_gather(Bw * world)
        for _ in range(3):
            gw(engw.generate(melw))
        wsteps = min(args.steps, 2)
        w_dev, w_e2e, _, _, _, _ = timed_passes(engw, melw, melw_host, idsw_host, wsteps, gw, sync_all, -1)
        tw = torch.tensor([w_dev, w_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        w_dev, w_e2e = tw.tolist()
        weak = {"value": round(AUDIO_SECONDS * Bw * world * wsteps / (w_dev / 1e3), 2), "unit": UNIT, "scaling": "weak",
                "batch_per_gpu": Bw, "global_batch": Bw * world, "steps": wsteps, "warmup": 3,
                "ms_per_step": round(w_dev / wsteps, 2), "e2e": round(AUDIO_SECONDS * Bw * world * wsteps / (w_e2e / 1e3), 2)}
        engw.close()
        eng = None
    del sd

    if rank == 0:
        H, d = cfg["decoder_attention_heads"], cfg["d_model"]
        # algorithmic bytes of one cross-attention launch: K and V of every (utterance, head) read once
        # (2 * 1500 * 64 elements) + q read + out written (X / L per utterance)
        rows = -(-B // args.streams) if args.streams > 1 else B      # the timed launches are those of sub-session 0
        xattn_bytes = rows * H * (2 * cfg["max_source_positions"] * 64 * es) + 2 * rows * d * es
        achieved = xattn_bytes / (xattn_ms / max(xattn_n, 1) * 1e-3) / 1e9 if xattn_n else None
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("decode_attn_cross_dram_bytes_per_launch")   # ncu capture at B = 256
            if traffic is not None and rows != 256:
                traffic = int(traffic * rows / 256)
        roofline = {"kernel": "decode_attn_kernel (cross-attention, 1 query x 1500 keys)", "bound": "hbm",
                    "achieved": round(achieved, 1) if achieved else None, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": round(achieved / peaks["hbm_gbs"], 4) if achieved else None, "traffic": traffic,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": xattn_bytes,
                    "avg_launch_us": round(xattn_ms / max(xattn_n, 1) * 1e3, 2), "launches_timed": xattn_n,
                    "timed": f"CUDA events on the launching stream around each launch of decode step {prof_step} of every "
                             "greedy loop inside the timed region (that step runs eagerly, the rest replay the CUDA graph)",
                    "share_of_step": round(xattn_ms / max(xattn_n, 1) * cfg["decoder_layers"] * (args.max_length - 1)
                                           * args.steps * args.streams / dev_ms, 4)}
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(dev_ms / args.steps, 2), "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": mel_host.numel() * 4,
                    "d2h_bytes_per_step": B * args.max_length * 4, "ms_per_step": round(e2e_ms / args.steps, 2)},
            "gpu_launches": int(launches), "streams_per_gpu": args.streams,
            "phases": {"encoder_ms": round(enc_ms, 1), "decode_ms": round(dec_ms, 1),
                       "decode_step_us": round(dec_ms / (args.max_length - 1) * 1e3, 1),
                       "note": "one extra identical pass after the timed regions, rank 0"},
            "clocks": clocks,
            "roofline": roofline,
        }
        # the whole job against its own roofline (per GPU; ranks are independent): north_star's "fraction of roofline"
        job = job_ceiling(cfg, B, es, peaks)
        job["frac"] = round(value / world / job["ceiling"], 4)
        line["job_roofline"] = job
        if probe is not None:
            line["parity_probe"] = probe
        if micro is not None:
            line["decode_step_us"] = micro
        if weak is not None:
            line["weak_scaling"] = weak
        if not args.no_cpu_baseline and world == 1:   # rank 0 at N = 1 only: at N > 1 the other ranks' host threads spin in NCCL
            r = cpu_reference_rtfx(args.size, args.max_length, 16, 128)
            line["cpu_baseline"] = {"value": round(r["value"], 3), "unit": UNIT, "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"]}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if eng is not None:
        eng.close()


if __name__ == "__main__":
    main()

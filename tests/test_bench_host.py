"""bench.py's host arithmetic (no GPU): the whole-job roofline reproduces SURVEY.md §8d's figures, and the JSON contract
fields of the reference arm are in place."""
import json
import subprocess
import sys

import bench
from conftest import ROOT
from whisper_trtllm_b200 import synthetic as synth

PEAKS = {"hbm_gbs": 6538.6, "bf16_tflops": 1619.6, "bf16_tflops_sustained": 1356.8}


def test_job_ceiling_reproduces_the_survey_figures():
    cfg = synth.make_config("medium.en", max_length=448)
    j = bench.job_ceiling(cfg, 256, 2, PEAKS)
    assert abs(j["flops_per_utterance"] / 1e9 - 1289.1) < 0.5          # 1138.1 encoder + 151.0 cross-K/V GFLOP
    assert abs(j["decode_bytes_per_step_mean"] / 1e9 - 44.2) < 0.3      # "~44 GB per step" at B = 256
    assert abs(j["encoder_floor_ms"] - 243) < 2 and abs(j["decode_floor_ms"] - 3021) < 15
    assert abs(j["ceiling"] - 2350) < 10                                # "~2350x RT ceiling per GPU"
    # per-rank batch of strong scaling: the weight stream weighs more, the ceiling drops
    assert bench.job_ceiling(cfg, 32, 2, PEAKS)["ceiling"] < j["ceiling"]
    # a peaks file without the sustained figure still works
    assert bench.job_ceiling(cfg, 256, 2, {"hbm_gbs": 6538.6, "bf16_tflops": 1619.6})["ceiling"] > j["ceiling"]


def test_workload_config_names_the_workload():
    class A:   # --batch 256: weak scaling, 256 utterances per GPU
        size, dtype, batch, global_batch, max_length = "medium.en", "bf16", 256, 256, 448
    c = bench.workload_config(A, 8)
    assert c["global_batch"] == 2048 and c["parallelism"] == "dp8" and "medium.en" in c["workload"] and "model" not in c
    assert bench.batch_plan(A, 8, 3) == (256, 256, 2048, "weak")

    class S:   # default: BASELINE.json configs[3], a global batch of 256 sharded over the ranks
        size, dtype, batch, global_batch, max_length = "medium.en", "bf16", 0, 256, 448
    for world, per in ((1, 256), (2, 128), (4, 64), (8, 32)):
        c = bench.workload_config(S, world)
        assert c["global_batch"] == 256 and c["batch_per_gpu"] == per and c["parallelism"] == f"dp{world}"
        assert [bench.batch_plan(S, world, r)[0] for r in range(world)] == [per] * world
        assert bench.batch_plan(S, world, 0)[3] == "strong"
    S.global_batch = 10                                  # ragged shards: earlier ranks take the remainder
    assert [bench.batch_plan(S, 4, r)[0] for r in range(4)] == [3, 3, 2, 2] and bench.batch_plan(S, 4, 0)[1] == 3


def test_cpu_sample_shrinks_with_the_number_of_passes():
    assert bench.pick_cpu_sample(6) == (16, 128)         # default bench run: the largest sample
    assert bench.pick_cpu_sample(25) in ((8, 64), (4, 128), (8, 128))
    assert bench.pick_cpu_sample(1000) == (4, 64)


def test_reference_arm_prints_one_contract_line():
    # tiny.en keeps the CPU sample short; the line has the reference-arm keys of the contract
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--size", "tiny.en", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    # ms_per_step is what actually RAN (the measured sample), never an extrapolation; the extrapolated job time is beside it
    cb = d["cpu_baseline"]
    assert abs(cb["measured_ms_per_sample"] - d["ms_per_step"]) < 1e-6 and cb["extrapolated_ms_per_sample"] >= cb["measured_ms_per_sample"]
    assert d["scaling"] == "strong" and d["config"]["global_batch"] == 256
    assert d["e2e"] == {"value": d["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}

"""bench.py's GPU arm cannot execute in a container without a GPU; its CONTROL FLOW can: tests/_bench_flow_worker.py runs it
with the device replaced by host stand-ins and the trainer by a stub that issues a training step's collective.  What this
holds: every rank issues the same number of steps although the ranks' host clocks disagree (the worker skews them by 7 %
per rank) and the ranks run at different speeds — a mismatch deadlocks, which is how the thermal-settle loop of commit
7c99aa7 ("until 1.2 s have passed" on each rank's own clock) hung the 8-GPU strong-scaling run, and this test hangs on that
bench.py in the same way; rank 0 prints exactly one contract line; nothing in the arm raises."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
WORKER = REPO / "tests" / "_bench_flow_worker.py"
ARGS = ["--steps", "6", "--warmup", "3", "--batch", "2", "--classes", "16"]


def _run(world, port, extra=()):
    env = dict(os.environ, REPO=str(REPO), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world))
    procs = [subprocess.Popen([sys.executable, str(WORKER), "--gpus", str(world), *ARGS, *extra],
                              env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              text=True, cwd=str(REPO)) for r in range(world)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-3000:]
    return outs


def _steps(err):
    return int([ln for ln in err.splitlines() if ln.startswith("rank ")][-1].split()[-1])


@pytest.mark.parametrize("world", [1, 2])
def test_gpu_arm_control_flow_without_a_gpu(world):
    outs = _run(world, 29650 + world, extra=("--no-cpu-baseline", "--no-eager-baseline"))
    lines = [ln for o, _ in outs for ln in o.splitlines() if ln.startswith("{")]
    assert len(lines) == 1                                   # rank 0 alone prints, once
    line = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in line, k
    assert line["n_gpus"] == world and line["steps"] == 6 and line["config"]["global_batch"] == 2 * world
    assert line["value"] > 0 and line["e2e"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 2 * 3 * 224 * 224 + 2 * 8
    assert line["config"]["thermal_settle_steps"] >= 4
    counts = [_steps(e) for _, e in outs]
    assert len(set(counts)) == 1, counts                     # every rank issued the same number of steps (collectives)
    # warm-up + settle + three timed regions (device-resident, end to end with its two lead-in steps, instrumented)
    assert counts[0] == 3 + line["config"]["thermal_settle_steps"] + 6 + 2 + 6 + 6


def test_gpu_arm_control_flow_strong_scaling_two_ranks():
    outs = _run(2, 29655, extra=("--config", "3", "--scaling", "strong", "--batch", "4", "--no-settle"))
    line = json.loads([ln for o, _ in outs for ln in o.splitlines() if ln.startswith("{")][0])
    assert line["scaling"] == "strong" and line["config"]["global_batch"] == 4 and line["config"]["thermal_settle_steps"] == 0
    assert len({_steps(e) for _, e in outs}) == 1


def test_gpu_arm_control_flow_eval():
    outs = _run(1, 29657, extra=("--eval",))
    line = json.loads([ln for o, _ in outs for ln in o.splitlines() if ln.startswith("{")][0])
    assert line["metric"].startswith("inference images/sec") and line["value"] > 0 and line["cpu_baseline"] is None

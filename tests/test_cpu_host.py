"""CPU-only checks: the C-ABI library loads and exports everything include/mvlpt_sm100.h declares, the host-side
logic (prompt index maps, config, optimiser schedule, data-parallel sharding over gloo) and loud failure without CUDA."""
import ctypes
import math
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

REPO = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from mvlpt_b200 import _lib
    from mvlpt_b200.build import build_lib
    build_lib()
    return _lib.lib()


def test_library_exports_every_declared_symbol(lib):
    from mvlpt_b200 import _abi
    protos = _abi.prototypes()
    assert len(protos) >= 24
    for name in protos:
        assert hasattr(lib, name), name
    assert lib.mvlpt_version() == 1
    assert lib.mvlpt_launch_count() == 0
    for must in ["mvlpt_gemm", "mvlpt_fmha_fwd", "mvlpt_fmha_bwd", "mvlpt_ln_fwd", "mvlpt_ln_bwd", "mvlpt_ce_fwd_bwd",
                 "mvlpt_text_assemble", "mvlpt_embed_assemble", "mvlpt_prompt_grad", "mvlpt_ctx_grad", "mvlpt_sgd",
                 "mvlpt_upt_fwd", "mvlpt_upt_bwd"]:
        assert must in protos, must


def test_sass_is_blackwell_native():
    """The linear kernel must be tcgen05 + TMA (UTC*MMA / UTMALDG in SASS), not the legacy HMMA path."""
    obj = REPO / "mvlpt_b200" / "csrc" / "_obj" / "gemm.o"
    if not obj.exists():
        pytest.skip("object files not kept")
    sass = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass


def test_no_gpu_is_an_error_not_a_fallback(lib):
    from mvlpt_b200 import _lib, ops
    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only container")
    assert lib.mvlpt_check_device(0) != 0
    assert lib.mvlpt_last_error()
    a = torch.zeros(8, 8, dtype=torch.half)
    with pytest.raises(_lib.MvlptError):
        ops.gemm(a, a, a.clone())
    from mvlpt_b200.trainers.mvlpt import MVLPT
    from mvlpt_b200.trainers.runtime import default_cfg
    with pytest.raises(_lib.MvlptError):
        MVLPT(default_cfg(), dm=object())


def test_product_never_imports_the_oracle():
    for p in (REPO / "mvlpt_b200").rglob("*.py"):
        text = p.read_text()
        assert "import oracle" not in text and "from oracle" not in text, p


def test_ctx_maps_reproduce_forward_coop_layouts():
    """slot / ctx_pos maps vs the oracle's literal restatement of forward_coop (trainers/mvlpt.py:455-510)."""
    from mvlpt_b200 import engine as E
    from oracle import mvlpt_oracle as O
    torch.manual_seed(0)
    C, Lt, n, d = 4, 20, 6, 8
    name_lens = [1, 3, 2, 5]
    emb = torch.randn(C, Lt, d)
    ctx = torch.randn(n, d)
    for position in ("end", "middle", "front"):
        ref = O.coop_prompts(emb, ctx, name_lens, n, position)
        slot, pos = E.build_ctx_maps(name_lens, n, Lt, position)
        base = E.rearrange_embedding(emb, name_lens, n, position)
        mine = base.clone()
        for c in range(C):
            for t in range(Lt):
                if slot[c, t] >= 0:
                    mine[c, t] = ctx[slot[c, t]]
            for j in range(n):
                assert torch.equal(ref[c, pos[c, j]], ctx[j])
        assert torch.equal(mine, ref), position


def test_cfg_yaml_and_opts_merge(tmp_path):
    from mvlpt_b200.trainers import runtime as R
    y = tmp_path / "vit_b16.yaml"
    y.write_text("DATALOADER:\n  TRAIN_X:\n    BATCH_SIZE: 32\nINPUT:\n  SIZE: (224, 224)\nOPTIM:\n  NAME: \"sgd\"\n  LR: 0.002\n"
                 "  MAX_EPOCH: 200\n  LR_SCHEDULER: \"cosine\"\n  WARMUP_EPOCH: 1\n  WARMUP_TYPE: \"constant\"\n"
                 "  WARMUP_CONS_LR: 1e-5\nMODEL:\n  BACKBONE:\n    NAME: \"ViT-B/16\"\n")
    cfg = R.merge_yaml(R.default_cfg(), str(y))
    assert cfg.INPUT.SIZE == (224, 224) and cfg.OPTIM.MAX_EPOCH == 200 and cfg.MODEL.BACKBONE.NAME == "ViT-B/16"
    R.merge_list(cfg, ["TRAINER.MVLPT.VPT.N_CTX", "8", "TRAINER.MVLPT.COOP.CLASS_TOKEN_POSITION", "middle",
                       "TRAINER.CUT_CONTEXTLEN", "True"])
    assert cfg.TRAINER.MVLPT.VPT.N_CTX == 8 and cfg.TRAINER.CUT_CONTEXTLEN is True
    assert cfg.TRAINER.MVLPT.COOP.CLASS_TOKEN_POSITION == "middle"


def test_lr_schedule_constant_warmup_then_cosine():
    from mvlpt_b200.trainers import runtime as R
    p = torch.nn.Parameter(torch.zeros(3))
    opt = R.PromptSGD([("p", p)], lr=0.002)
    sch = R.ConstantWarmupCosine(opt, max_epoch=200, warmup_epoch=1, cons_lr=1e-5)
    assert opt.lr == 1e-5
    sch.step()
    assert opt.lr == pytest.approx(0.002)
    sch.step()
    assert opt.lr == pytest.approx(0.002 * 0.5 * (1 + math.cos(math.pi * 1 / 200)))
    ref = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=0.002)
    cos = torch.optim.lr_scheduler.CosineAnnealingLR(ref, 200.0)
    for _ in range(10):
        ref.step()
        cos.step()
        sch.step()
    assert opt.lr == pytest.approx(cos.get_last_lr()[0] if False else 0.002 * 0.5 * (1 + math.cos(math.pi * 11 / 200)))


def test_flop_accounting_matches_survey_numbers():
    from mvlpt_b200 import accounting, synth
    a = synth.ARCHS["ViT-B/16"]
    assert accounting.flops_step(a, 256, 100, 77, 0, 16) / 1e12 == pytest.approx(10.20, abs=0.05)   # SURVEY §8d cfg 2
    assert accounting.flops_step(a, 256, 2193, 77, 8, 0) / 1e12 == pytest.approx(32.1, abs=0.3)     # cfg 3
    from oracle import mvlpt_oracle as O
    assert O.flops_step(a, 256, 1000, 77, 8, 16) == accounting.flops_step(a, 256, 1000, 77, 8, 16)


_WORKER = r"""
import os, sys, torch
sys.path.insert(0, os.environ["REPO"])
from mvlpt_b200.trainers.runtime import DataParallelGroup
dp = DataParallelGroup.from_env("gloo")
assert dp.world == 2
# image shard: contiguous, balanced, covers everything once
lo, hi = dp.shard(7)
cover = torch.zeros(7); cover[lo:hi] = 1
dp.all_reduce_sum(cover)
assert torch.equal(cover, torch.ones(7)), cover
# gradient exchange: each rank holds the gradient of its half batch (already divided by the GLOBAL batch);
# SUM all-reduce must equal the full-batch gradient
torch.manual_seed(0)
per_sample = torch.randn(8, 5)
full = per_sample.sum(0) / 8
mine = per_sample[dp.rank * 4:(dp.rank + 1) * 4].sum(0) / 8
dp.all_reduce_sum(mine)
assert torch.allclose(mine, full, atol=1e-6)
t = torch.tensor([float(dp.rank + 1)], dtype=torch.float64)
dp.all_reduce_max(t)
assert float(t) == 2.0
# bench.py's thermal-settle count: ranks that measured different step times must still run the SAME number of steps
# (each step issues a gradient all-reduce; a count taken from a rank's own clock deadlocks the job)
import bench
n = bench.settle_steps(dp, 0.010 * (dp.rank + 1), torch.device("cpu"))
agreed = torch.tensor([float(n), -float(n)], dtype=torch.float64)
dp.all_reduce_max(agreed)
assert n == 60 and agreed.tolist() == [60.0, -60.0], (n, agreed)
assert bench.settle_steps(dp, 1e-9, torch.device("cpu")) == 512 and bench.settle_steps(dp, 5.0, torch.device("cpu")) == 1
# class-sharded text tower (SURVEY.md 8e): per-rank class ranges tile [0, C); padded all-gather of the features;
# reduce-scatter of their gradients gives every rank the SUM over ranks for its own classes
C, e = 7, 3
ranges = [dp.shard(C, r) for r in range(dp.world)]
assert ranges[0][0] == 0 and ranges[-1][1] == C and all(ranges[i][1] == ranges[i + 1][0] for i in range(dp.world - 1))
cmax = max(b - a for a, b in ranges)
feats = torch.arange(C * e, dtype=torch.float32).view(C, e)           # what a single process would compute
c0, c1 = ranges[dp.rank]
send = torch.zeros(cmax, e); send[:c1 - c0] = feats[c0:c1]
gathered = torch.zeros(dp.world, cmax, e)
dp.all_gather_into(gathered, send)
full = torch.cat([gathered[r, :b - a] for r, (a, b) in enumerate(ranges)])
assert torch.equal(full, feats)
g_local = torch.full((C, e), float(dp.rank + 1))                      # d(loss)/d(features) from this rank's images
g_send = torch.zeros(dp.world, cmax, e)
for r, (a, b) in enumerate(ranges):
    g_send[r, :b - a] = g_local[a:b]
g_recv = torch.zeros(cmax, e)
dp.reduce_scatter_sum(g_recv, g_send)
assert torch.equal(g_recv[:c1 - c0], torch.full((c1 - c0, e), 3.0)), g_recv
dp.barrier()
print("rank", dp.rank, "ok")
"""


def test_data_parallel_group_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    env = dict(os.environ, REPO=str(REPO), MASTER_ADDR="127.0.0.1", MASTER_PORT="29631", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_bench_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--batch", "4", "--classes", "4", "--ctx-len", "32"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    import json
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "cpu_baseline", "e2e", "config", "higher_is_better"):
        assert k in line
    # the unmodified reference modules when oracle/_ref was built (oracle/build_ref.py), else the oracle port
    from oracle import build_ref
    kind = "reference" if build_ref.available() else "port"
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == kind and line["value"] > 0
    assert line["config"]["same_config"] is True  # 4 of 4 images: the sample is the whole (tiny) batch


@pytest.mark.skipif(not os.path.isdir("/root/reference/configs/trainers/MVLPT"), reason="reference checkout not present")
def test_reference_yaml_files_load_unchanged_and_select_the_transform_stack():
    """the reference's own trainer YAMLs drop in through the cfg shim (container-only test)."""
    import glob
    from mvlpt_b200.trainers import runtime as R
    from mvlpt_b200.input_pipeline import build_transform
    files = sorted(glob.glob("/root/reference/configs/trainers/MVLPT/vit_*.yaml"))
    assert files
    for f in files:
        cfg = R.merge_yaml(R.default_cfg(), f)
        side = 336 if cfg.MODEL.BACKBONE.NAME.endswith("@336px") else 224
        assert cfg.MODEL.BACKBONE.NAME.startswith("ViT-") and tuple(cfg.INPUT.SIZE) == (side, side)
        assert cfg.OPTIM.LR_SCHEDULER == "cosine" and cfg.OPTIM.WARMUP_TYPE == "constant"
        t = build_transform(cfg, True)
        assert t.mode == "train" and t.flip_p == 0.5 and abs(t.mean[0] - 0.48145466) < 1e-7 and t.size == (side, side)
        assert build_transform(cfg, False).mode == "test"
    with pytest.raises(NotImplementedError):  # Dassl's own default interpolation is bilinear: not what MVLPT runs with
        build_transform(R.default_cfg(), True)


def test_cli_config_precedence_matches_train_py():
    """train.py:171-191: defaults <- dataset YAML <- trainer YAML <- named arguments <- free KEY VALUE options; the flag
    set is the reference's (train.py:223-293)."""
    from mvlpt_b200 import train as T
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False) as f:
        f.write("OPTIM:\n  LR: 0.5\n  MAX_EPOCH: 7\nMODEL:\n  BACKBONE:\n    NAME: \"ViT-B/16\"\nINPUT:\n  SIZE: (224, 224)\n")
        path = f.name
    a = T.build_parser().parse_args(["--trainer", "MVLPT", "--config-file", path, "--output-dir", "/tmp/x", "--seed", "3",
                                     "--multi-task", "--multi-task-label_pertask", "--dataset-coop", "--cut-contextlen",
                                     "--backbone", "ViT-B/32", "--resume", "/tmp/r", "--shots", "5",
                                     "--multi-task-evalkey", "dtd",
                                     "OPTIM.MAX_EPOCH", "9", "TRAINER.MVLPT.COOP.N_CTX", "16", "TEST.FINAL_MODEL", "best_val"])
    c = T.setup_cfg(a)
    assert c.OPTIM.LR == 0.5 and c.OPTIM.MAX_EPOCH == 9          # YAML, then overridden by the free option
    assert c.MODEL.BACKBONE.NAME == "ViT-B/32"                   # named argument beats the YAML
    assert c.TRAINER.NAME == "MVLPT" and c.OUTPUT_DIR == "/tmp/x" and c.RESUME == "/tmp/r" and c.SEED == 3
    assert c.DATASET.MULTITASK and c.DATASET.MULTITASK_LABEL_PERTASK and c.DATASET.COOP and c.TRAINER.CUT_CONTEXTLEN
    assert c.DATASET.NUM_SHOTS == 5 and c.DATASET.MULTITASK_EVALKEY == "dtd" and c.DATASET.RANDOM_SEED_SAMPLING == 3
    assert c.TRAINER.MVLPT.COOP.N_CTX == 16 and c.TEST.FINAL_MODEL == "best_val" and tuple(c.INPUT.SIZE) == (224, 224)
    # defaults of extend_cfg (train.py:130-169)
    d = T.setup_cfg(T.build_parser().parse_args([]))
    M = d.TRAINER.MVLPT
    assert (M.PREC, M.PROJECT_METHOD, M.PROJECT_DIM, M.VPT.N_CTX, M.VPT.DEEP, M.COOP.CLASS_TOKEN_POSITION) == \
        ("fp16", "transformer", 128, 0, True, "middle")
    assert not d.TRAINER.CUT_CONTEXTLEN and d.TRAINER.ACT_CKPT == 1 and d.DATASET.MULTITASK_EVALKEY == "average"


def test_trainer_registry_resolves_mvlpt():
    from mvlpt_b200.trainers.runtime import TRAINER_REGISTRY
    from mvlpt_b200.trainers.mvlpt import MVLPT
    assert TRAINER_REGISTRY.get("MVLPT") is MVLPT
    with pytest.raises(KeyError):
        TRAINER_REGISTRY.get("NoSuchTrainer")


@pytest.mark.skipif(not os.path.isdir("/root/reference/configs/trainers/MVLPT"), reason="reference checkout not present")
def test_cli_reads_the_reference_yaml_unchanged():
    from mvlpt_b200 import train as T
    a = T.build_parser().parse_args(["--trainer", "MVLPT", "--config-file",
                                     "/root/reference/configs/trainers/MVLPT/vit_b16.yaml", "--dataset-coop", "--multi-task",
                                     "TRAINER.MVLPT.VPT.N_CTX", "8", "TRAINER.MVLPT.COOP.N_CTX", "8",
                                     "TRAINER.MVLPT.COOP.CLASS_TOKEN_POSITION", "middle", "TRAINER.MVLPT.COOP.CSC", "False",
                                     "TEST.NO_TEST", "False", "TEST.FINAL_MODEL", "best_val", "TRAINER.CUT_CONTEXTLEN", "True"])
    c = T.setup_cfg(a)   # the option list of scripts/mvlpt/main_mt_coopdata_cut.sh:30-47
    assert c.OPTIM.LR == 0.002 and c.OPTIM.MAX_EPOCH == 200 and c.OPTIM.WARMUP_CONS_LR == 1e-5
    assert c.DATALOADER.TRAIN_X.BATCH_SIZE == 32 and c.DATALOADER.TEST.BATCH_SIZE == 100
    assert c.TRAINER.MVLPT.COOP.CSC is False and c.TEST.NO_TEST is False and c.TRAINER.CUT_CONTEXTLEN is True


def test_no_undefined_names_in_gpu_only_code():
    """bench.py's GPU arm and the trainers cannot execute without a GPU; a scope-aware check keeps NameErrors out of them."""
    r = subprocess.run([sys.executable, str(REPO / "tools" / "lint_names.py")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout


def test_bench_watchdog_ends_a_hung_run():
    """A rank stuck in a collective must not hold its GPU until the caller's limit: the watchdog exits 124; a finished run
    is not affected."""
    code = "import bench, time; bench.arm_watchdog(0.3); time.sleep(20); print('not reached')"
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, cwd=str(REPO))
    assert r.returncode == 124 and "watchdog" in r.stderr and "not reached" not in r.stdout
    r = subprocess.run([sys.executable, "-c", "import bench; bench.arm_watchdog(60); print('done')"],
                       capture_output=True, text=True, timeout=120, cwd=str(REPO))
    assert r.returncode == 0 and "done" in r.stdout


def test_loss_summary_reads_lazily_and_survives_every_way_of_copying_a_dict():
    """forward_backward's return value (trainers/mvlpt.py:940-946 keys) waits for ITS step's copy on first use only, and
    dict(summary) / {**summary} / json / MetricMeter-style update all see the numbers, never the unread placeholders."""
    import json
    import torch
    from mvlpt_b200.trainers.mvlpt import LossSummary

    class Event:
        waits = 0

        def synchronize(self):
            Event.waits += 1

    host = torch.tensor([1.5, 75.0])
    want = {"loss": 1.5, "acc": 75.0, "num_tasks": 3}
    mk = lambda: LossSummary(host, Event(), {"num_tasks": 3})  # noqa: E731
    s = mk()
    assert Event.waits == 0 and "loss" in s and len(s) == 3 and Event.waits == 0   # nothing read yet
    assert s["loss"] == 1.5 and s["acc"] == 75.0 and Event.waits == 1 and s.get("num_tasks") == 3 and Event.waits == 1
    for copy in (dict(mk()), {**mk()}, mk().copy(), json.loads(json.dumps(mk())), dict(mk().items()),
                 {k: mk()[k] for k in mk()}):
        assert copy == want
    meter = {}
    meter.update(mk())
    assert meter == want and mk() == want and list(mk().keys()) == list(want) and list(mk().values()) == list(want.values())

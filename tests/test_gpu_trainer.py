"""The MVLPT trainer mirror on the GPU: forward_backward contract, SGD trajectory vs the CPU oracle, checkpoints."""
from types import SimpleNamespace as NS

import pytest
import torch

pytestmark = pytest.mark.gpu


def _trainer(name, prec="fp32", lr=0.002):
    from mvlpt_b200.trainers.mvlpt import MVLPT
    from mvlpt_b200.trainers.runtime import default_cfg
    from tests.helpers import case_inputs
    fx, case, arch, sd, image, pp, upt = case_inputs(name)
    cfg = default_cfg()
    T = cfg.TRAINER.MVLPT
    T.PREC = prec
    T.PROJECT_METHOD = case.get("project_method", "identity")
    T.PROJECT_DIM = case.get("project_dim", 128)
    T.VPT.N_CTX, T.VPT.DEEP = case.get("vpt_n_ctx", 0), case.get("vpt_deep", False)
    T.COOP.N_CTX, T.COOP.CSC = case.get("coop_n_ctx", 0), case.get("csc", False)
    T.COOP.CLASS_TOKEN_POSITION = case.get("position", "end")
    cfg.TRAINER.CUT_CONTEXTLEN = case.get("cut", False)
    cfg.INPUT.SIZE = (arch["image_resolution"],) * 2
    cfg.DATASET.COOP = True
    cfg.OPTIM.LR = lr
    cfg.OPTIM.WARMUP_EPOCH = 0
    dm = NS(dataset=NS(classnames=fx["names"]), lab2cname=dict(enumerate(fx["names"])), num_classes=case["C"])
    tr = MVLPT(cfg, dm=dm, clip_state_dict=sd, tokenized_prompts=fx["tokenized_prompts"], name_lens=fx["name_lens"])
    tr.model.prompt_learner.load_state_dict(pp, strict=False)
    return tr, fx, case, sd, image, pp, upt


@pytest.mark.parametrize("name", ["tiny_coop_end", "tiny_vpt_deep", "tiny_upt_identity"])
def test_three_sgd_steps_follow_the_oracle(name):
    from oracle import mvlpt_oracle as O
    from tests.helpers import oracle_kwargs, rel_err
    tr, fx, case, sd, image, pp, upt = _trainer(name, "fp32", lr=0.5)  # large lr so the trajectory is visible
    tr.num_batches = 100
    keys = list(pp)
    params = [pp[k].clone() for k in keys]
    bufs = [None] * len(keys)
    kw = oracle_kwargs(fx, case, sd, upt)
    batch = {"img": image, "label": fx["label"], "domain": torch.zeros(len(fx["label"]), dtype=torch.long)}
    for step in range(3):
        o_logits, o_loss, o_grads = O.train_step(image, fx["label"], sd, dict(zip(keys, params)), **kw)
        bufs = O.sgd_step(params, [o_grads[k] for k in keys], bufs, lr=0.5)
        s = tr.forward_backward(batch)
        assert set(s) == {"loss", "acc"}
        assert abs(s["loss"] - float(o_loss)) < 5e-3 * max(1.0, float(o_loss))
        acc = 100.0 * float((o_logits.argmax(-1) == fx["label"]).float().mean())
        assert abs(s["acc"] - acc) < 1e-3
    mine = dict(tr.model.prompt_learner.named_parameters())
    for k, p in zip(keys, params):
        assert rel_err(mine[k].detach().float().cpu().reshape(p.shape), p) < 5e-3, k


def test_checkpoint_roundtrip_and_reference_key_names(tmp_path):
    tr, fx, case, sd, image, pp, upt = _trainer("tiny_upt_identity")
    sdict = tr.model.prompt_learner.state_dict()
    assert {"ctx", "vpt_embeddings", "vpt_embeddings_deep", "token_prefix", "token_suffix"} <= set(sdict)
    tr.save_model(3, str(tmp_path), is_best=True)  # Dassl's layout: 0-based epoch 3 -> model.pth.tar-4, epoch field 4
    ck = torch.load(tmp_path / "prompt_learner" / "model.pth.tar-4", weights_only=False)
    assert ck["epoch"] == 4 and {"state_dict", "optimizer", "scheduler", "val_result"} <= set(ck)
    assert (tmp_path / "prompt_learner" / "checkpoint").read_text().strip() == "model.pth.tar-4"
    with torch.no_grad():
        tr.model.prompt_learner.ctx.zero_()
    tr.load_model(str(tmp_path), epoch=4)
    assert torch.equal(tr.model.prompt_learner.ctx.detach().cpu().float(), pp["ctx"])
    tr.load_model(str(tmp_path))  # model-best.pth.tar


def test_eval_path_and_text_feature_cache():
    tr, fx, case, sd, image, pp, upt = _trainer("tiny_vpt_deep")
    out1 = tr.model_inference(image.cuda())
    out2 = tr.model_inference(image.cuda())  # second call reuses the cached (constant) text features
    assert torch.equal(out1, out2)
    from tests.helpers import rel_err
    assert rel_err(out1.float().cpu(), fx["logits"]) < 3e-3
    loader = [{"img": image, "label": fx["label"], "domain": torch.zeros(len(fx["label"]), dtype=torch.long)}]
    tr.test_loader = loader
    res = tr.test()
    acc = 100.0 * float((fx["logits"].argmax(-1) == fx["label"]).float().mean())
    assert abs(res - acc) < 1e-6 and tr.last_test_results["results"] == {"accuracy": res}


def test_eval_multitask_elevater_metrics_and_held_text_features():
    """trainers/mvlpt.py:989-1088, ELEVATER branch: per-task class slices, the task's own metric, MULTITASK_EVALKEY;
    with trainable context the text tower runs once for the whole loader."""
    from types import SimpleNamespace as NS
    from mvlpt_b200 import _lib
    from mvlpt_b200.trainers import metrics as MT
    tr, fx, case, sd, image, pp, upt = _trainer("tiny_coop_end")
    C = tr.model.prompt_learner.n_cls
    half = C // 2
    B = image.shape[0]
    tr.multi_task = True
    tr.cfg.DATASET.COOP = False
    tr.cfg.DATASET.MULTITASK_EVALKEY = "average"
    tr.dm._id2task = {0: "a", 1: "b"}
    tr.dm._task_class_idx = {"a": (0, half), "b": (half, C)}
    tr.dm._metric_name = {"a": "accuracy", "b": "11point_mAP"}
    tr.dm._metric = {}
    task = torch.tensor([i % 2 for i in range(B)])
    lab = torch.zeros(B, C)
    for i in range(B):
        lo, hi = (0, half) if task[i] == 0 else (half, C)
        lab[i, lo + (int(fx["label"][i]) % (hi - lo))] = 1
    h = max(1, B // 2)
    loader = [{0: image[:h], 1: lab[:h], 3: task[:h]}, {0: image[h:], 1: lab[h:], 3: task[h:]}]
    tr.test_loader = loader
    tr.model.cache_text_features = True
    n0 = _lib.launch_count()
    res = tr.test()
    with torch.no_grad():
        logits = tr.model(image.cuda()).float().cpu()
    want_a = MT.accuracy(lab[task == 0][:, :half].argmax(1).numpy(), logits[task == 0][:, :half].numpy())
    want_b = MT.map_11_points(lab[task == 1][:, half:].numpy(), logits[task == 1][:, half:].numpy())
    assert abs(tr.last_test_results["per_task"]["a"] - want_a) < 1e-9
    assert abs(tr.last_test_results["per_task"]["b"] - want_b) < 1e-9
    assert abs(res - (want_a + want_b) / 2) < 1e-9


def test_run_epoch_lookahead_staging_changes_nothing():
    """run_epoch stages batch i+1 (pinned host -> device on the copy stream, two buffers) before step i is enqueued: the
    parameters after an epoch are bit-identical to plain forward_backward calls on the same host batches."""
    from mvlpt_b200 import synth
    finals = []
    for mode in ("plain", "epoch"):
        tr, fx, case, sd, image, pp, upt = _trainer("tiny_vpt_deep", "fp16", lr=0.1)
        B, C = 4, case["C"]
        batches = []
        for i in range(5):
            g = torch.Generator().manual_seed(40 + i)
            batches.append({"img": synth.synth_images(B, image.shape[-1], seed=60 + i).half().pin_memory(),
                            "label": torch.randint(0, C, (B,), generator=g).pin_memory(),
                            "domain": torch.zeros(B, dtype=torch.long)})
        if mode == "plain":
            tr.num_batches = len(batches)
            for tr.batch_idx, b in enumerate(batches):
                tr.forward_backward(b)
        else:
            tr.train_loader_x = batches
            out = tr.run_epoch()
            assert set(out) == {"loss", "acc"} and tr.batch_idx == len(batches) - 1
        torch.cuda.synchronize()
        finals.append({k: p.detach().clone() for k, p in tr.model.prompt_learner.named_parameters()})
    for k in finals[0]:
        assert torch.equal(finals[0][k], finals[1][k]), k
    # a staged batch is accepted wherever a batch is
    staged = tr.stage_batch(batches[0])
    assert staged["img"].is_cuda and staged["label"].is_cuda and hasattr(staged["img"], "_mvlpt_ready")
    assert tr.stage_batch({"img": batches[0]["img"].cuda(), "label": batches[0]["label"]})["img"].is_cuda


@pytest.mark.parametrize("name", ["tiny_coop_end", "tiny_vpt_deep"])
def test_eval_ragged_last_batch_reads_fresh_text_features(name):
    """ELEVATER/Dassl test loaders keep the last (smaller) batch.  The held text features belong to the label space, not
    to a batch size: the logits test() collects for every batch equal per-batch model() calls after the parameters moved."""
    from tests.helpers import rel_err
    from mvlpt_b200 import synth
    tr, fx, case, sd, image, pp, upt = _trainer(name)
    C = case["C"]
    imgs = synth.synth_images(7, image.shape[-1], seed=11)
    labs = torch.arange(7) % C
    mk = lambda a, b: {"img": imgs[a:b], "label": labs[a:b], "domain": torch.zeros(b - a, dtype=torch.long)}
    tr.test_loader = [mk(0, 3), mk(3, 6), mk(6, 7)]  # ragged: 3, 3, 1
    tr.test()  # first evaluation: buffers for B=3 and B=1 now hold these parameters' features
    with torch.no_grad():  # move the trainable prompts, as an epoch of training would
        for p in tr.model.prompt_learner.parameters():
            p.add_(0.05 * torch.randn_like(p))
    tr.model._txt_cache_valid = False
    collected = []
    orig = tr.model_inference
    tr.model_inference = lambda inp, task=None: collected.append(orig(inp, task=task).float().cpu()) or collected[-1].cuda()
    tr.test()
    tr.model_inference = orig
    assert [c.shape[0] for c in collected] == [3, 3, 1]
    tr.model.hold_text_features(False)
    for (a, b), got in zip([(0, 3), (3, 6), (6, 7)], collected):
        with torch.no_grad():
            want = tr.model(imgs[a:b].cuda()).float().cpu()
        assert rel_err(got, want) < 1e-6, (a, b)


def test_fp32_eval_outputs_do_not_alias_the_engine_buffer():
    """PREC=fp32: CustomCLIP.forward must hand out its own tensor — two same-size batches kept by the caller (as test()
    does) stay different."""
    from mvlpt_b200 import synth
    tr, fx, case, sd, image, pp, upt = _trainer("tiny_vpt_deep", "fp32")
    a = synth.synth_images(3, image.shape[-1], seed=21).cuda()
    b = synth.synth_images(3, image.shape[-1], seed=22).cuda()
    oa = tr.model_inference(a)
    keep = oa.clone()
    ob = tr.model_inference(b)
    assert oa.data_ptr() != ob.data_ptr()
    assert torch.equal(oa, keep) and not torch.equal(oa, ob)
    labs = torch.tensor([0, 1, 2])
    tr.test_loader = [{"img": a.cpu(), "label": labs, "domain": torch.zeros(3, dtype=torch.long)},
                      {"img": b.cpu(), "label": labs, "domain": torch.zeros(3, dtype=torch.long)}]
    res = tr.test()
    want = 100.0 * float((torch.cat([oa, ob]).argmax(1).cpu() == torch.cat([labs, labs])).float().mean())
    assert abs(res - want) < 1e-6


def test_cli_trains_resumes_and_evaluates(tmp_path, capsys):
    """python -m mvlpt_b200.train (the reference's train.py flags, train.py:171-295) on the reference-shaped config keys
    with the synthetic stand-ins: 2 epochs with per-epoch validation (TEST.FINAL_MODEL best_val) -> checkpoints in Dassl's
    layout; --resume continues at epoch 2 and ends bit-identical to an uninterrupted 3-epoch run; --eval-only
    --model-dir --load-epoch reproduces the final test result."""
    from mvlpt_b200 import train as T

    def run(*flags, opts=()):
        # named flags first, free KEY VALUE options last (argparse.REMAINDER, exactly like the reference's train.py)
        return T.main(T.build_parser().parse_args(list(flags) + list(base_opts) + list(opts)))

    common = ["--trainer", "MVLPT", "--synthetic", "--synthetic-classes", "6", "--synthetic-batches", "3", "--seed", "1",
              "--backbone", "tiny", "--dataset-coop"]
    base_opts = ["TRAINER.MVLPT.VPT.N_CTX", "4", "TRAINER.MVLPT.COOP.N_CTX", "4", "TRAINER.MVLPT.PROJECT_METHOD", "identity",
                 "TRAINER.MVLPT.COOP.CLASS_TOKEN_POSITION", "end", "INPUT.SIZE", "(64, 64)",
                 "DATALOADER.TRAIN_X.BATCH_SIZE", "4", "DATALOADER.TEST.BATCH_SIZE", "5", "OPTIM.LR", "0.05",
                 "OPTIM.WARMUP_EPOCH", "0", "TEST.FINAL_MODEL", "best_val", "OPTIM.MAX_EPOCH", "3"]
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    full = run(*common, "--output-dir", a)
    assert (tmp_path / "a" / "prompt_learner" / "model.pth.tar-3").exists()
    assert (tmp_path / "a" / "prompt_learner" / "model-best.pth.tar").exists()
    assert isinstance(full.final_result, float)
    # interrupt the same 3-epoch schedule after 2 epochs: checkpoint every epoch, then pretend epoch 3 never happened
    run(*common, "--output-dir", b, opts=["TRAIN.CHECKPOINT_FREQ", "1", "TEST.NO_TEST", "True"])
    import os
    os.remove(tmp_path / "b" / "prompt_learner" / "model.pth.tar-3")
    (tmp_path / "b" / "prompt_learner" / "checkpoint").write_text("model.pth.tar-2\n")
    res = run(*common, "--output-dir", b, "--resume", b, opts=["TEST.NO_TEST", "True"])
    assert res.start_epoch == 2
    assert "Previous epoch: 2" in capsys.readouterr().out
    ck_a = torch.load(tmp_path / "a" / "prompt_learner" / "model.pth.tar-3", weights_only=False)
    ck_b = torch.load(tmp_path / "b" / "prompt_learner" / "model.pth.tar-3", weights_only=False)
    for k, v in ck_a["state_dict"].items():
        assert torch.equal(v, ck_b["state_dict"][k]), k
    ev = run(*common, "--eval-only", "--model-dir", a, "--load-epoch", "3")
    with torch.no_grad():
        want = ev.test()
    last = run(*common, "--output-dir", str(tmp_path / "c"), opts=["TEST.FINAL_MODEL", "last_step"])
    assert abs(last.final_result - want) < 1e-6

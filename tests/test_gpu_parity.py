"""GPU parity: the CUDA path (through the C ABI, via the reference-API mirror) against the CPU oracle on the same
seeded inputs and against the golden fixtures generated from the reference itself (oracle/gen_golden.py).

Bars (normwise-max relative error  max|a-b| / max|b|  against the reference's fp32 run; BASELINE.json north star:
1e-3 fp16-relative; SURVEY.md App. A: gradients 5e-3, and never worse than the reference's own fp16 path):

  logits     <= 1e-3, or — where fp16 operand rounding makes that unreachable — <= the error of the reference's OWN fp16 run
             of the same step (`ref16_logits` in the fixture) with 50 % slack for the spread between two independent fp16
             evaluations; never above 2.5e-3
  gradients  <= 5e-3, or <= 1.5 x the reference's own fp16 error of that gradient (`ref16_grads`)
  argmax     bit-exact on every sample whose fp32 top-2 margin exceeds 4 fp16 ulps.

The measured numbers per fixture are committed in profiles/r02_parity_report.json (tools/gpu_parity_report.py): full-size
fixtures sit at 0.2-1.9e-3 (logits) against 1.4-7.0e-3 for the reference's fp16 path.  CoCoOp's first meta-net layer
is the one ill-conditioned gradient: the oracle ITSELF moves it by 1.2e-2 when only the token embeddings are rounded to
fp16 (what PREC="fp16" mandates, trainers/mvlpt.py:307), so it is held to 8 x its ref16 error instead.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

LOGIT_TOL, LOGIT_CAP = 1e-3, 2.5e-3
GRAD_TOL = 5e-3
REF16_SLACK = 1.5


def logit_bar(ref16_err=None):
    return LOGIT_TOL if ref16_err is None else min(LOGIT_CAP, max(LOGIT_TOL, REF16_SLACK * ref16_err))


def grad_bar(name, ref16_err=None):
    if ref16_err is None:
        return GRAD_TOL
    if name.startswith("meta_net.linear1"):
        # ill-conditioned (a sum over classes of terms that cancel, behind a ReLU): the fp32 oracle itself moves it by
        # 1.2e-2 when only the token embeddings are rounded to fp16, as PREC="fp16" mandates (measured)
        return max(GRAD_TOL, 8.0 * ref16_err)
    return max(GRAD_TOL, REF16_SLACK * ref16_err)


TINY = ["tiny_coop_end", "tiny_coop_middle_cut", "tiny_coop_front_csc", "tiny_vpt_shallow", "tiny_vpt_deep",
        "tiny_vpt_deep_taskmask_soft", "tiny_upt_identity", "tiny_upt_transformer", "tiny_cocoop", "tiny_cocoop_vpt_deep",
        "tiny_vpt_deep_project", "tiny_vpt_shallow_project_coop"]
FULL = ["b16_coop_end", "b16_vpt_deep", "b16_upt_transformer", "b32_coop_cfg1", "l14_coop_end", "b16_cocoop",
        "b16_vpt_deep_project", "l14_vpt_deep", "l14_336_vpt_deep", "b16_coop_c1000_cut", "b16_vpt_deep_11task"]


def _run(name, prec):
    from tests.helpers import build_custom_clip, rel_err
    model, fx, case, sd, image, pp, upt = build_custom_clip(name, prec)
    img = image.cuda()
    if prec == "fp16":
        img = img.half()
    loss_rows, pred, grads = model.loss_and_grads(img, fx["label"].cuda(), fx["task"])
    torch.cuda.synchronize()
    logits = model.last_logits(img.shape[0]).float().cpu()
    return model, fx, case, sd, image, pp, upt, logits, loss_rows.cpu(), pred.cpu(), {k: g.cpu() for k, g in grads.items()}


def _ref16_errors(fx):
    """Errors of the reference's own fp16 run against its fp32 run: (logits, {gradient name: error})."""
    from tests.helpers import rel_err
    if "ref16_logits" not in fx:
        return None, {}
    return (rel_err(fx["ref16_logits"], fx["logits"]),
            {k: rel_err(g.reshape(fx["grads"][k].shape), fx["grads"][k]) for k, g in fx["ref16_grads"].items()})


def _check_against(fx_logits, fx_loss, fx_grads, margin, logits, loss_rows, pred, grads, ref16=(None, {})):
    from tests.helpers import rel_err
    r16_logits, r16_grads = ref16
    le = rel_err(logits, fx_logits)
    assert le <= logit_bar(r16_logits), f"logits {le:.2e} > bar {logit_bar(r16_logits):.2e} (reference fp16: {r16_logits})"
    assert abs(float(loss_rows.mean()) - float(fx_loss)) <= 5e-3 * max(1.0, abs(float(fx_loss)))
    ulp = 9.8e-4 * fx_logits.abs().max()
    safe = margin > 4 * ulp
    assert torch.equal(logits.argmax(-1)[safe], fx_logits.argmax(-1)[safe])
    assert torch.equal(pred.long()[safe], fx_logits.argmax(-1)[safe])
    for k, g in fx_grads.items():
        assert k in grads, f"missing gradient {k}"
        ge, bar = rel_err(grads[k].reshape(g.shape), g), grad_bar(k, r16_grads.get(k))
        assert ge <= bar, f"gradient {k}: {ge:.2e} > bar {bar:.2e} (reference fp16: {r16_grads.get(k)})"


@pytest.mark.parametrize("prec", ["fp32", "fp16"])
@pytest.mark.parametrize("name", TINY + FULL)
def test_step_matches_reference_golden(name, prec):
    """logits, loss, argmax and every prompt gradient vs the fixture produced by the reference's own CustomCLIP."""
    model, fx, case, sd, image, pp, upt, logits, loss_rows, pred, grads = _run(name, prec)
    _check_against(fx["logits"], fx["loss"], fx["grads"], fx["top2_margin"], logits, loss_rows, pred, grads,
                   ref16=_ref16_errors(fx))


@pytest.mark.parametrize("name", TINY)
def test_step_matches_oracle_on_fresh_inputs(name):
    """Same comparison against the oracle run live on differently seeded images/labels (not the fixture's)."""
    from oracle import mvlpt_oracle as O
    from mvlpt_b200 import synth
    from tests.helpers import build_custom_clip, oracle_kwargs
    model, fx, case, sd, image, pp, upt = build_custom_clip(name, "fp32")
    B = 5
    image = synth.synth_images(B, synth.ARCHS[case["arch"]]["image_resolution"], seed=77)
    g = torch.Generator().manual_seed(123)
    kw = oracle_kwargs(fx, case, sd, upt)
    task = None
    if fx["task"] is not None:
        task = torch.randint(0, len(case["tasks"]), (B,), generator=g)
        kw["task"] = task
    if fx["label"].dim() > 1:
        label = torch.zeros(B, case["C"])
        label[torch.arange(B), torch.randint(0, case["C"], (B,), generator=g)] = 1.0
        label[torch.arange(B), torch.randint(0, case["C"], (B,), generator=g)] = 1.0
    else:
        label = torch.randint(0, case["C"], (B,), generator=g)
    o_logits, o_loss, o_grads = O.train_step(image, label, sd, pp, **kw)
    loss_rows, pred, grads = model.loss_and_grads(image.cuda(), label.cuda(), task)
    torch.cuda.synchronize()
    logits = model.last_logits(B).float().cpu()
    top2 = o_logits.topk(2, dim=-1).values
    # other images, same configuration: the fixture's ref16 errors are the yard-stick of that configuration
    _check_against(o_logits, o_loss, o_grads, top2[:, 0] - top2[:, 1], logits, loss_rows.cpu(), pred.cpu(),
                   {k: v.cpu() for k, v in grads.items()}, ref16=_ref16_errors(fx))


@pytest.mark.parametrize("name", ["tiny_coop_end", "tiny_vpt_deep", "tiny_upt_identity"])
def test_autograd_path_equals_fused_path(name):
    """CustomCLIP.forward + F.cross_entropy + .backward() (how the reference trainer drives it) gives the same
    gradients as the fused loss_and_grads call."""
    import torch.nn.functional as F
    from tests.helpers import build_custom_clip, rel_err
    model, fx, case, sd, image, pp, upt = build_custom_clip(name, "fp32")
    img, lab = image.cuda(), fx["label"].cuda()
    _, _, grads = model.loss_and_grads(img, lab, fx["task"])
    fused = {k: v.clone() for k, v in grads.items()}
    out = model(img, task=fx["task"])
    F.cross_entropy(out.float(), lab).backward()
    for k, p in model.prompt_learner.named_parameters():
        if k in fused:
            assert rel_err(p.grad.float().reshape(fused[k].shape), fused[k]) < 2e-3, k


def test_encoders_match_golden_features():
    """ImageEncoder / TextEncoder called directly (reference signatures) reproduce the reference features."""
    from tests.helpers import build_custom_clip, rel_err
    model, fx, case, sd, image, pp, upt = build_custom_clip("tiny_vpt_deep", "fp32")
    f = model.image_encoder(image.cuda())
    assert rel_err(f.float().cpu(), fx["image_features"]) < 5e-4
    model, fx, case, sd, image, pp, upt = build_custom_clip("tiny_coop_middle_cut", "fp32")
    prompts = model.prompt_learner.forward_coop()
    assert rel_err(prompts.float().cpu(), fx["prompts"]) < 1e-3
    t = model.text_encoder(prompts, model.tokenized_prompts)
    assert rel_err(t.float().cpu(), fx["text_features"]) < 1e-3


def test_full_size_properties():
    """BASELINE-size batch (ViT-B/16, B=64 here to bound test time): size-independent properties —
    (1) batch independence: logits of image i do not depend on the other images in the batch;
    (2) gradient linearity: the prompt gradient of a batch equals the weighted sum over two half batches;
    (3) the frozen CLIP weights are bit-identical before and after a step."""
    from mvlpt_b200 import synth
    from tests.helpers import build_custom_clip, rel_err
    model, fx, case, sd, image, pp, upt = build_custom_clip("b16_vpt_deep", "fp16")
    B, C = 64, case["C"]
    img = synth.synth_images(B, 224, seed=5).half().cuda()
    lab = torch.randint(0, C, (B,), generator=torch.Generator().manual_seed(1)).cuda()
    tower = model.image_encoder.tower(img.device)
    w_before = [b.w_qkv.clone() for b in tower.blocks[:2]]
    _, _, g_full = model.loss_and_grads(img, lab)
    g_full = {k: v.clone() for k, v in g_full.items()}
    logits_full = model.last_logits(B).clone()
    _, _, g_a = model.loss_and_grads(img[:32], lab[:32], global_batch=B)
    g_a = {k: v.clone() for k, v in g_a.items()}
    logits_a = model.last_logits(32).clone()
    _, _, g_b = model.loss_and_grads(img[32:], lab[32:], global_batch=B)
    assert torch.equal(logits_full[:32], logits_a)
    for k in g_full:
        assert rel_err(g_a[k] + g_b[k], g_full[k]) < 2e-3, k
    for b, w in zip(tower.blocks[:2], w_before):
        assert torch.equal(b.w_qkv, w)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tiny_coop_end", "tiny_coop_front_csc", "b16_coop_end"])
def test_causal_cut_of_the_text_tower_changes_nothing(name, monkeypatch):
    """Rows behind the last EOT cannot reach an EOT row through the causal mask (clip/model.py:324-330): running the text
    tower on the first max(eot)+1 rows must give the logits and the context gradients of all context_length rows."""
    from tests.helpers import build_custom_clip, rel_err
    out = {}
    # same attention tiling on both sides (packing several short sequences to a tile changes the fp32 summation order
    # inside the tensor core, nothing else; it is compared separately below)
    monkeypatch.setenv("MVLPT_FMHA_PACK", "0")
    for cut in ("1", "0"):
        monkeypatch.setenv("MVLPT_TEXT_CAUSAL_CUT", cut)
        model, fx, case, sd, image, pp, upt = build_custom_clip(name, "fp16", device="cuda")
        pl = model.prompt_learner
        assert pl.context_len == fx["tokenized_prompts"].shape[1]
        assert (pl.kernel_len < pl.context_len) == (cut == "1")
        loss_rows, pred, grads = model.loss_and_grads(image.cuda().half(), fx["label"].cuda(), fx["task"])
        torch.cuda.synchronize()
        out[cut] = (model.last_logits(image.shape[0]).float().clone(), {k: v.float().clone() for k, v in grads.items()},
                    model.prompt_learner.forward_coop().float().clone())
    assert torch.equal(out["1"][0], out["0"][0]), "logits differ"
    for k in out["0"][1]:
        assert rel_err(out["1"][1][k], out["0"][1][k]) < 1e-6, k
    assert torch.equal(out["1"][2], out["0"][2]), "forward_coop (API parity, full length) differs"
    # packed short sequences (the default) against one sequence per tile: same numbers up to summation order
    monkeypatch.setenv("MVLPT_FMHA_PACK", "1")
    monkeypatch.setenv("MVLPT_TEXT_CAUSAL_CUT", "1")
    model, fx, case, sd, image, pp, upt = build_custom_clip(name, "fp16", device="cuda")
    loss_rows, pred, grads = model.loss_and_grads(image.cuda().half(), fx["label"].cuda(), fx["task"])
    torch.cuda.synchronize()
    # (two equally valid fp16 evaluations: they differ by what each differs from the fp32 reference, measured 4e-3 on the
    # ViT-B/16 context gradient)
    assert rel_err(model.last_logits(image.shape[0]).float(), out["1"][0]) < LOGIT_CAP
    for k in out["1"][1]:
        assert rel_err(grads[k].float(), out["1"][1][k]) < 1e-2, k


@pytest.mark.parametrize("prec", ["fp32", "fp16"])
@pytest.mark.parametrize("name", ["tiny_vpt_deep_dropout", "tiny_vpt_shallow_project_dropout"])
def test_vpt_dropout_training_step_matches_oracle_on_the_kernels_masks(name, prec):
    """vpt_dropout in training mode (trainers/mvlpt.py:76,165,425).  The reference draws its masks from torch's RNG; the
    oracle is pinned to it with those draws GIVEN (tests/test_oracle_golden.py, fixtures `drop_keep`).  Here the kernels
    draw their own counter-based masks; mvlpt_dropout_keep returns them and the oracle replays the step with them."""
    from oracle import mvlpt_oracle as O
    from mvlpt_b200 import ops
    from tests.helpers import build_custom_clip, oracle_kwargs, rel_err
    model, fx, case, sd, image, pp, upt = build_custom_clip(name, prec)
    pl = model.prompt_learner
    p = case["vpt_dropout"]
    assert pl.training and pl.vpt_dropout.p == p
    pl.drop_seed_override = 20240229
    img = image.cuda().half() if prec == "fp16" else image.cuda()
    B, v = image.shape[0], case["vpt_n_ctx"]
    d = model.image_encoder.tower(img.device).d
    n_slabs = 1 + (pl.vpt_embeddings_deep.shape[0] if case.get("vpt_deep") else 0)
    loss_rows, pred, grads = model.loss_and_grads(img, fx["label"].cuda(), fx["task"])
    logits = model.last_logits(B).float().cpu()
    keep = []
    for slab in range(n_slabs):
        k = torch.empty(B, v, d, dtype=torch.uint8, device="cuda")
        ops.dropout_keep(k, B, v, d, p, pl.drop_seed_override, slab)
        keep.append(k.bool().cpu())
    rate = torch.stack(keep).float().mean().item()
    n = torch.stack(keep).numel()
    assert abs(rate - (1 - p)) < 5 * (p * (1 - p) / n) ** 0.5 + 1e-4, rate
    if n_slabs > 1:
        assert not torch.equal(keep[0], keep[1])
    kw = oracle_kwargs(fx, case, sd, upt)
    kw["drop_keep"] = keep
    o_logits, o_loss, o_grads = O.train_step(image, fx["label"], sd, pp, **kw)
    top2 = o_logits.topk(2, dim=-1).values
    base = name.replace("_dropout", "") if name != "tiny_vpt_shallow_project_dropout" else "tiny_vpt_shallow_project_coop"
    from tests.conftest import load_golden
    _check_against(o_logits, o_loss, o_grads, top2[:, 0] - top2[:, 1], logits, loss_rows.cpu(), pred.cpu(),
                   {k: g.cpu() for k, g in grads.items()}, ref16=(_ref16_errors(load_golden(base))[0], {}))
    # another seed: another mask, other logits; eval mode: no dropout at all (nn.Dropout semantics)
    pl.drop_seed_override = 7
    model.loss_and_grads(img, fx["label"].cuda(), fx["task"])
    assert not torch.equal(model.last_logits(B).float().cpu(), logits)
    pl.eval()
    kw["drop_p"], kw["drop_keep"] = 0.0, None
    e_logits, _, _ = O.train_step(image, fx["label"], sd, pp, **kw)
    model.loss_and_grads(img, fx["label"].cuda(), fx["task"])
    assert rel_err(model.last_logits(B).float().cpu(), e_logits) <= LOGIT_CAP

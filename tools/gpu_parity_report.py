"""Print parity numbers of the CUDA path against the golden fixtures / oracle for every case (run under gpurun)."""
import sys, time, traceback
sys.path.insert(0, ".")
import torch
from tests.helpers import build_custom_clip, rel_err

names = sys.argv[1:] or ["tiny_coop_end", "tiny_coop_middle_cut", "tiny_coop_front_csc", "tiny_vpt_shallow", "tiny_vpt_deep",
                         "tiny_vpt_deep_taskmask_soft", "tiny_upt_identity", "b16_coop_end", "b16_vpt_deep", "b32_coop_cfg1",
                         "l14_coop_end", "tiny_cocoop", "tiny_cocoop_vpt_deep", "tiny_upt_transformer", "b16_upt_transformer",
                         "tiny_vpt_deep_project", "tiny_vpt_shallow_project_coop", "b16_cocoop", "b16_vpt_deep_project",
                         "l14_vpt_deep"]
bad = 0
for prec in ("fp32", "fp16"):
    for name in names:
        try:
            model, fx, case, sd, image, pp, upt = build_custom_clip(name, prec)
            img = image.cuda()
            if prec == "fp16":
                img = img.half()
            loss_rows, pred, grads = model.loss_and_grads(img, fx["label"].cuda(), fx["task"])
            torch.cuda.synchronize()
            logits = model.last_logits(img.shape[0]).float().cpu()
            loss = float(loss_rows.mean())
            le = rel_err(logits, fx["logits"])
            same = bool(torch.equal(logits.argmax(-1), fx["logits"].argmax(-1)))
            ge = {k: rel_err(g.cpu().reshape(fx["grads"][k].shape), fx["grads"][k]) for k, g in grads.items()}
            print(f"{prec} {name}: logits_rel={le:.2e} loss={loss:.6f} (ref {float(fx['loss']):.6f}) argmax_same={same} "
                  f"min_margin={float(fx['top2_margin'].min()):.3f} grads_rel={ {k: f'{v:.2e}' for k, v in ge.items()} }", flush=True)
            if le > 3e-3 or max(v for k, v in ge.items() if not k.startswith('meta_net.linear1')) > 2e-2:
                bad += 1
            del model
            torch.cuda.empty_cache()
        except Exception:
            bad += 1
            print(f"{prec} {name}: EXCEPTION"); traceback.print_exc()
print("BAD" if bad else "ALL OK", bad)

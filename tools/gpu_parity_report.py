"""Per-fixture parity numbers of the CUDA path (run under gpurun): python tools/gpu_parity_report.py [out.json] [cases..]

For every golden fixture x {fp32, fp16} mode: normwise-max relative error (max|a-b| / max|b|) of the
logits, the loss, every prompt gradient and — for the tiny fixtures, which carry them — the image / text features,
against the REFERENCE's fp32 run; next to each, the error of the reference's OWN fp16 run (`ref16_*` in the fixture,
oracle/gen_golden.py) against the same fp32 run.  Written as JSON (committed as profiles/rNN_parity_report.json): the
parity tests' bars are read off this table, not guessed.
"""
import json
import sys
import traceback

sys.path.insert(0, ".")
import torch

from tests.conftest import GOLDEN
from tests.helpers import build_custom_clip, rel_err

args = sys.argv[1:]
out_path = args.pop(0) if args and args[0].endswith(".json") else None
names = args or sorted(p.stem for p in GOLDEN.glob("*.pt") if p.stem not in ("metrics", "preprocess")
                       and "dropout" not in p.stem)
report, bad = {}, 0
for name in names:
    for prec in ("fp32", "fp16"):
        key = f"{name}/{prec}"
        try:
            model, fx, case, sd, image, pp, upt = build_custom_clip(name, prec)
            img = image.cuda().half() if prec == "fp16" else image.cuda()
            loss_rows, pred, grads = model.loss_and_grads(img, fx["label"].cuda(), fx["task"])
            torch.cuda.synchronize()
            B = img.shape[0]
            logits = model.last_logits(B).float().cpu()
            r = dict(
                logits=rel_err(logits, fx["logits"]),
                loss_abs=abs(float(loss_rows.mean()) - float(fx["loss"])),
                argmax_equal=bool(torch.equal(logits.argmax(-1), fx["logits"].argmax(-1))),
                min_top2_margin=float(fx["top2_margin"].min()), max_abs_logit=float(fx["logits"].abs().max()),
                grads={k: rel_err(g.cpu().reshape(fx["grads"][k].shape), fx["grads"][k]) for k, g in grads.items()
                       if k in fx["grads"]})
            if "ref16_logits" in fx:
                r["ref16_logits"] = rel_err(fx["ref16_logits"], fx["logits"])
                r["ref16_argmax_equal"] = bool(torch.equal(fx["ref16_logits"].argmax(-1), fx["logits"].argmax(-1)))
                r["ref16_grads"] = {k: rel_err(g.reshape(fx["grads"][k].shape), fx["grads"][k])
                                    for k, g in fx["ref16_grads"].items()}
            if "image_features" in fx:
                pl = model.prompt_learner
                with torch.no_grad():
                    ctx, vpt, deep = pl.forward_mvlpt_proj()
                    f = model.image_encoder(img, vpt, deep)
                    r["image_features"] = rel_err(f.float().cpu(), fx["image_features"])
                    if "text_features" in fx:
                        t = model.text_encoder(pl.forward_coop(ctx), model.tokenized_prompts)
                        r["text_features"] = rel_err(t.float().cpu(), fx["text_features"])
            report[key] = r
            gmax = max(r["grads"].values()) if r["grads"] else 0.0
            g16 = max(r.get("ref16_grads", {"-": float("nan")}).values())
            print(f"{key}: logits {r['logits']:.2e} (ref16 {r.get('ref16_logits', float('nan')):.2e}) "
                  f"grads max {gmax:.2e} (ref16 {g16:.2e}) argmax_equal={r['argmax_equal']} "
                  f"feat img {r.get('image_features', float('nan')):.2e} txt {r.get('text_features', float('nan')):.2e}",
                  flush=True)
            del model
            torch.cuda.empty_cache()
        except Exception:
            bad += 1
            report[key] = {"exception": traceback.format_exc()}
            print(f"{key}: EXCEPTION")
            traceback.print_exc()
if out_path:
    with open(out_path, "w") as f:
        json.dump(report, f, indent=1, sort_keys=True)
print("EXCEPTIONS" if bad else "DONE", bad)

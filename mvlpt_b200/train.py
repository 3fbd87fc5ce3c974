"""Entry point mirroring the reference's train.py (argparse -> cfg -> build_trainer -> train / test; train.py:171-295).

    python -m mvlpt_b200.train --trainer MVLPT --config-file configs/trainers/MVLPT/vit_b16.yaml \\
        --output-dir out/ [--resume DIR] [--eval-only --model-dir DIR --load-epoch N] [--no-train] \\
        [--multi-task] [--multi-task-label_pertask] [--dataset-coop] [--cut-contextlen] KEY VALUE ...

Same flags, same precedence (dataset YAML < trainer YAML < named arguments < free `KEY VALUE` options; train.py:171-191)
and the same config keys (`extend_cfg`, train.py:105-169) as the reference, over the attribute-style cfg of
trainers/runtime.py — Dassl / yacs are not installable offline.

What the reference takes from the network or the file system has to be handed in:
  * CLIP weights: `MODEL.BACKBONE.PATH <state-dict file>` (the reference downloads them, clip/clip.py:29-70);
  * data: `--data-manager pkg.module:factory` names a callable `factory(cfg) -> DataManager-like object` (the attribute
    surface of SURVEY.md App. F); the reference builds it from datasets on disk (trainers/mvlpt.py:883-908, out of scope);
  * `--synthetic` supplies both from mvlpt_b200.synth (random-init CLIP of the configured backbone, random class-name
    token ids, synthetic batches): the offline smoke path of this CLI, also what the tests drive.
"""
from __future__ import annotations

import argparse
import importlib
import random
import sys

import numpy as np
import torch


def build_parser() -> argparse.ArgumentParser:
    """The reference's flags (train.py:223-293), plus --data-manager / --synthetic* (see module docstring)."""
    p = argparse.ArgumentParser(prog="python -m mvlpt_b200.train")
    p.add_argument("--root", type=str, default="", help="path to dataset")
    p.add_argument("--output-dir", type=str, default="", help="output directory")
    p.add_argument("--resume", type=str, default="", help="checkpoint directory (from which the training resumes)")
    p.add_argument("--seed", type=int, default=-1, help="only positive value enables a fixed seed")
    p.add_argument("--source-domains", type=str, nargs="+", help="source domains for DA/DG")
    p.add_argument("--target-domains", type=str, nargs="+", help="target domains for DA/DG")
    p.add_argument("--transforms", type=str, nargs="+", help="data augmentation methods")
    p.add_argument("--config-file", type=str, default="", help="path to config file")
    p.add_argument("--dataset-config-file", type=str, default="", help="path to config file for dataset setup")
    p.add_argument("--dataset", type=str, default="", help="name of task")
    p.add_argument("--shots", type=int, help="few shot")
    p.add_argument("--trainer", type=str, default="", help="name of trainer")
    p.add_argument("--backbone", type=str, default="", help="name of CNN backbone")
    p.add_argument("--head", type=str, default="", help="name of head")
    p.add_argument("--eval-only", action="store_true", help="evaluation only")
    p.add_argument("--model-dir", type=str, default="", help="load model from this directory for eval-only mode")
    p.add_argument("--load-epoch", type=int, help="load model weights at this epoch for evaluation")
    p.add_argument("--no-train", action="store_true", help="do not call trainer.train()")
    p.add_argument("--multi-task", action="store_true")
    p.add_argument("--multi-task-label_pertask", action="store_true")
    p.add_argument("--multi-task-evalkey", type=str, default="average")
    p.add_argument("--dataset-coop", action="store_true")
    p.add_argument("--cut-contextlen", action="store_true")
    p.add_argument("--act-ckpt", type=int, default=1)
    # offline stand-ins for what the reference downloads / reads from disk
    p.add_argument("--data-manager", type=str, default="", help="pkg.module:factory, factory(cfg) -> DataManager-like")
    p.add_argument("--synthetic", action="store_true", help="synthetic CLIP weights, class tokens and batches")
    p.add_argument("--synthetic-classes", type=int, default=10)
    p.add_argument("--synthetic-batches", type=int, default=4)
    p.add_argument("opts", default=None, nargs=argparse.REMAINDER, help="modify config options using the command-line")
    return p


def reset_cfg(cfg, args):
    """train.py:48-103, key for key."""
    D = cfg.DATASET
    if args.root:
        D.ROOT = args.root
    if args.output_dir:
        cfg.OUTPUT_DIR = args.output_dir
    if args.resume:
        cfg.RESUME = args.resume
    if args.seed:
        cfg.SEED = args.seed
        D.RANDOM_SEED_SAMPLING = args.seed
    if args.source_domains:
        D.SOURCE_DOMAINS = args.source_domains
    if args.target_domains:
        D.TARGET_DOMAINS = args.target_domains
    if args.transforms:
        cfg.INPUT.TRANSFORMS = args.transforms
    if args.trainer:
        cfg.TRAINER.NAME = args.trainer
    if args.backbone:
        cfg.MODEL.BACKBONE.NAME = args.backbone
    if args.head:
        if not hasattr(cfg.MODEL, "HEAD"):
            from types import SimpleNamespace as NS
            cfg.MODEL.HEAD = NS()
        cfg.MODEL.HEAD.NAME = args.head
    if args.dataset:
        D.DATASET = args.dataset
    if args.shots:
        D.NUM_SAMPLES_PER_CLASS = args.shots
        D.NUM_SHOTS = args.shots
    if args.multi_task:
        D.MULTITASK = args.multi_task
    if args.multi_task_label_pertask:
        D.MULTITASK_LABEL_PERTASK = args.multi_task_label_pertask
    if args.dataset_coop:
        D.COOP = args.dataset_coop
    if args.cut_contextlen:
        cfg.TRAINER.CUT_CONTEXTLEN = args.cut_contextlen
    if args.act_ckpt:
        cfg.TRAINER.ACT_CKPT = args.act_ckpt
    if args.multi_task_evalkey != "average":
        D.MULTITASK_EVALKEY = args.multi_task_evalkey


def setup_cfg(args):
    """train.py:171-191: defaults (+ extend_cfg keys) <- dataset YAML <- trainer YAML <- named arguments <- opts."""
    from .trainers import runtime as R
    cfg = R.default_cfg()
    if args.dataset_config_file:
        R.merge_yaml(cfg, args.dataset_config_file)
    if args.config_file:
        R.merge_yaml(cfg, args.config_file)
    reset_cfg(cfg, args)
    opts = [o for o in (args.opts or []) if o != "--"]
    R.merge_list(cfg, opts)
    return cfg


def set_random_seed(seed: int):
    """dassl.utils.set_random_seed (upstream): python, numpy and torch generators."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def build_trainer(cfg, **kw):
    """dassl.engine.build_trainer: look `cfg.TRAINER.NAME` up in the trainer registry."""
    from .trainers.runtime import TRAINER_REGISTRY
    from .trainers import mvlpt  # noqa: F401  (registers MVLPT)
    name = getattr(cfg.TRAINER, "NAME", "") or "MVLPT"
    return TRAINER_REGISTRY.get(name)(cfg, **kw)


def _synthetic_inputs(cfg, args):
    from . import synth
    arch = cfg.MODEL.BACKBONE.NAME
    T = cfg.TRAINER.MVLPT
    n_text = T.COCOOP.N_CTX or T.COOP.N_CTX
    C = args.synthetic_classes
    toks, name_lens = synth.synth_token_ids(C, n_text, context_length=synth.ARCHS[arch]["context_length"], seed=3)
    dm = synth.SyntheticDataManager(arch, C, train_batches=args.synthetic_batches,
                                    batch_size=cfg.DATALOADER.TRAIN_X.BATCH_SIZE,
                                    test_images=2 * cfg.DATALOADER.TEST.BATCH_SIZE + 3,
                                    test_batch_size=cfg.DATALOADER.TEST.BATCH_SIZE, seed=max(cfg.SEED, 0),
                                    half=T.PREC == "fp16")
    cfg.DATASET.COOP = True  # the synthetic batches are in the CoOp-data format
    return dict(dm=dm, clip_state_dict=synth.synth_clip_state_dict(arch, seed=0), tokenized_prompts=toks,
                name_lens=name_lens)


def main(args):
    """train.py:194-219."""
    cfg = setup_cfg(args)
    if cfg.SEED >= 0:
        print("Setting fixed seed: {}".format(cfg.SEED))
        set_random_seed(cfg.SEED)
    kw = {}
    if args.synthetic:
        kw = _synthetic_inputs(cfg, args)
    elif args.data_manager:
        mod, _, fn = args.data_manager.partition(":")
        kw = dict(dm=getattr(importlib.import_module(mod), fn)(cfg))
    trainer = build_trainer(cfg, **kw)
    if args.eval_only:
        trainer.load_model(args.model_dir, epoch=args.load_epoch)
        res = trainer.test()
        print("=> result\n* {}: {:.2f}".format(list(trainer.last_test_results["results"])[0], res))
        return trainer
    if args.model_dir:
        trainer.load_model(args.model_dir)
    if not args.no_train:
        trainer.train()
    return trainer


if __name__ == "__main__":
    main(build_parser().parse_args())

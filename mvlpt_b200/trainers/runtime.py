"""Host-side runtime the MVLPT trainer needs around the hot path, written against the call sites the reference
makes into Dassl — Dassl itself is not installable offline and is not part of this build:

  * `default_cfg()`            — attribute-style config with the keys/defaults of train.py:105-169 (`extend_cfg`) plus
                                 the few Dassl keys trainers/mvlpt.py reads; `merge_yaml()` overlays a reference YAML
                                 (configs/trainers/MVLPT/*.yaml) unchanged.
  * `PromptSGD`                — torch.optim.SGD as Dassl's build_optimizer configures it for trainers/mvlpt.py:869
                                 (momentum 0.9, weight decay 5e-4, dampening 0, no Nesterov), executed by the
                                 mvlpt_sgd kernel straight from the engine's flat fp32 gradient buffer.
  * `ConstantWarmupCosine`     — Dassl's build_lr_scheduler for LR_SCHEDULER=cosine, WARMUP_TYPE=constant
                                 (configs/trainers/MVLPT/vit_b16.yaml:15-22): per-EPOCH schedule.
  * `DataParallelGroup`        — one process per GPU over torch.distributed (replaces nn.DataParallel,
                                 trainers/mvlpt.py:877-880): SUM all-reduce of the flat prompt-gradient buffer.
"""
from __future__ import annotations

import math
import os
from types import SimpleNamespace as NS
from typing import Dict, Iterable, List, Optional, Tuple

import torch


# ---------------------------------------------------------------------------------------------------- config
def default_cfg() -> NS:
    """Defaults of the keys the hot path reads (train.py:118-169; Dassl defaults for OPTIM)."""
    return NS(
        TRAINER=NS(
            NAME="",
            MVLPT=NS(PREC="fp16", PROJECT_METHOD="transformer", PROJECT_DIM=128,
                     VPT=NS(N_CTX=0, CSC=False, CTX_INIT="", DROPOUT=0.0, PROJECT=-1, DEEP=True),
                     COOP=NS(N_CTX=0, CSC=False, CTX_INIT="", CLASS_TOKEN_POSITION="middle"),
                     COCOOP=NS(N_CTX=0, CTX_INIT="", PREC="fp16")),
            CUT_CONTEXTLEN=False, ACT_CKPT=1),
        # Dassl's INPUT defaults; every MVLPT YAML overrides INTERPOLATION / PIXEL_* / TRANSFORMS
        # (configs/trainers/MVLPT/vit_b16.yaml:8-13)
        INPUT=NS(SIZE=(224, 224), INTERPOLATION="bilinear", TRANSFORMS=(), PIXEL_MEAN=[0.485, 0.456, 0.406],
                 PIXEL_STD=[0.229, 0.224, 0.225], RRCROP_SCALE=(0.08, 1.0)),
        MODEL=NS(BACKBONE=NS(NAME="ViT-B/16", PATH=""), INIT_WEIGHTS=""),
        OPTIM=NS(NAME="sgd", LR=0.002, MAX_EPOCH=200, LR_SCHEDULER="cosine", WARMUP_EPOCH=1, WARMUP_TYPE="constant",
                 WARMUP_CONS_LR=1e-5, MOMENTUM=0.9, WEIGHT_DECAY=5e-4, SGD_DAMPNING=0, SGD_NESTEROV=False),
        DATASET=NS(COOP=False, MULTITASK=False, MULTITASK_LABEL_PERTASK=False, MULTITASK_EVALKEY="average", NAME="",
                   DATASET="", ROOT="", NUM_SHOTS=-1, NUM_SAMPLES_PER_CLASS=20, RANDOM_SEED_SAMPLING=1,
                   SUBSAMPLE_CLASSES="all", SOURCE_DOMAINS=(), TARGET_DOMAINS=()),
        DATALOADER=NS(TRAIN_X=NS(BATCH_SIZE=32), TEST=NS(BATCH_SIZE=100), NUM_WORKERS=8),
        TEST=NS(SPLIT="test", FINAL_MODEL="last_step", NO_TEST=False),
        TRAIN=NS(PRINT_FREQ=5, CHECKPOINT_FREQ=0),
        OUTPUT_DIR="", RESUME="", VERBOSE=False, USE_CUDA=True, SEED=-1,
    )


class Registry:
    """dassl.engine.TRAINER_REGISTRY's surface (upstream: a fvcore-style name -> class table): `register()` is a class
    decorator, `get(name)` looks a trainer up (train.py:208 build_trainer -> TRAINER_REGISTRY.get(cfg.TRAINER.NAME))."""

    def __init__(self, name: str):
        self._name, self._obj = name, {}

    def register(self, obj=None, force: bool = False):
        def deco(cls):
            if cls.__name__ in self._obj and not force:
                raise KeyError(f"{cls.__name__} is already registered in {self._name}")
            self._obj[cls.__name__] = cls
            return cls
        return deco if obj is None else deco(obj)

    def get(self, name: str):
        if name not in self._obj:
            raise KeyError(f"{name!r} is not registered in {self._name}; available: {sorted(self._obj)}")
        return self._obj[name]

    def registered_names(self):
        return list(self._obj)


TRAINER_REGISTRY = Registry("TRAINER")


def _merge(node: NS, d: dict, path: str = ""):
    for k, v in d.items():
        if isinstance(v, dict):
            if not hasattr(node, k):
                setattr(node, k, NS())
            _merge(getattr(node, k), v, f"{path}{k}.")
        else:
            if isinstance(v, str):
                # yacs' _decode_cfg_value: strings that are Python literals become them ("(224, 224)" -> tuple,
                # "1e-5" -> float — PyYAML's YAML-1.1 resolver leaves that one a string)
                import ast
                try:
                    v = ast.literal_eval(v)
                except (ValueError, SyntaxError):
                    pass
            setattr(node, k, v)


def merge_yaml(cfg: NS, path: str) -> NS:
    """Overlay a reference YAML file (yacs semantics: nested keys override)."""
    import yaml
    with open(path) as f:
        _merge(cfg, yaml.safe_load(f) or {})
    return cfg


def merge_list(cfg: NS, opts: Iterable) -> NS:
    """`KEY.SUB VALUE` pairs as train.py passes them after the named options (train.py:185-186)."""
    import ast
    opts = list(opts)
    assert len(opts) % 2 == 0, "opts must be KEY VALUE pairs"
    for k, v in zip(opts[0::2], opts[1::2]):
        node = cfg
        parts = k.split(".")
        for p in parts[:-1]:
            if not hasattr(node, p):
                setattr(node, p, NS())
            node = getattr(node, p)
        if isinstance(v, str):
            try:
                v = ast.literal_eval(v)
            except (ValueError, SyntaxError):
                pass
        setattr(node, parts[-1], v)
    return cfg


# ---------------------------------------------------------------------------------------------------- optimiser
class PromptSGD:
    """SGD with momentum over the prompt tensors.  `step(flat_grad)` consumes the engine's flat fp32 gradient buffer
    (CustomCLIP.grad_buffer(), named_parameters() order) with one mvlpt_sgd launch per tensor."""

    def __init__(self, named_params: List[Tuple[str, torch.nn.Parameter]], lr: float, momentum: float = 0.9,
                 weight_decay: float = 5e-4):
        self.named_params = [(n, p) for n, p in named_params if p.requires_grad]
        self.param_groups = [dict(lr=float(lr), initial_lr=float(lr), momentum=float(momentum),
                                  weight_decay=float(weight_decay), params=[p for _, p in self.named_params])]
        self.bufs: Dict[str, torch.Tensor] = {}
        self.steps = 0

    @property
    def lr(self) -> float:
        return self.param_groups[0]["lr"]

    def zero_grad(self, set_to_none: bool = True):
        for _, p in self.named_params:
            p.grad = None

    def step(self, flat_grad: torch.Tensor):
        from .. import ops
        g = self.param_groups[0]
        off = 0
        for n, p in self.named_params:
            k = p.numel()
            if n not in self.bufs:
                self.bufs[n] = torch.zeros_like(p.data)
            ops.sgd(p.data, self.bufs[n], flat_grad[off:off + k], g["lr"], g["momentum"], g["weight_decay"],
                    first_step=self.steps == 0)
            off += k
        self.steps += 1

    def state_dict(self):
        return dict(steps=self.steps, lr=self.lr, bufs={k: v.clone() for k, v in self.bufs.items()})

    def load_state_dict(self, sd):
        self.steps = sd["steps"]
        self.param_groups[0]["lr"] = sd["lr"]
        self.bufs = {k: v.clone() for k, v in sd["bufs"].items()}


class ConstantWarmupCosine:
    """lr(epoch) = WARMUP_CONS_LR for epoch < WARMUP_EPOCH, then base·½(1+cos(π·(epoch−warmup)/(MAX_EPOCH−warmup)))
    — Dassl's ConstantWarmupScheduler wrapping CosineAnnealingLR(T_max=MAX_EPOCH) hands the cosine its own epoch count,
    restarted after the warm-up."""

    def __init__(self, optim: PromptSGD, max_epoch: int, warmup_epoch: int = 1, cons_lr: float = 1e-5):
        self.optim, self.max_epoch, self.warmup, self.cons_lr = optim, int(max_epoch), int(warmup_epoch), float(cons_lr)
        self.base = optim.param_groups[0]["initial_lr"]
        self.last_epoch = 0
        self._apply()

    def lr_at(self, epoch: int) -> float:
        if epoch < self.warmup:
            return self.cons_lr
        return self.base * 0.5 * (1.0 + math.cos(math.pi * (epoch - self.warmup) / max(1, self.max_epoch)))

    def _apply(self):
        self.optim.param_groups[0]["lr"] = self.lr_at(self.last_epoch)

    def step(self):
        self.last_epoch += 1
        self._apply()

    def get_last_lr(self):
        return [self.optim.lr]

    def state_dict(self):
        return dict(last_epoch=self.last_epoch)

    def load_state_dict(self, sd):
        self.last_epoch = int(sd["last_epoch"])
        self._apply()


def build_optimizer(named_params, optim_cfg) -> PromptSGD:
    if getattr(optim_cfg, "NAME", "sgd") != "sgd":
        raise NotImplementedError("only OPTIM.NAME='sgd' (every reference MVLPT config) has a kernel")
    return PromptSGD(named_params, optim_cfg.LR, getattr(optim_cfg, "MOMENTUM", 0.9),
                     getattr(optim_cfg, "WEIGHT_DECAY", 5e-4))


def build_lr_scheduler(optim: PromptSGD, optim_cfg) -> ConstantWarmupCosine:
    if getattr(optim_cfg, "LR_SCHEDULER", "cosine") != "cosine":
        raise NotImplementedError("only LR_SCHEDULER='cosine' is implemented")
    warm = getattr(optim_cfg, "WARMUP_EPOCH", 0) if getattr(optim_cfg, "WARMUP_TYPE", "constant") == "constant" else 0
    return ConstantWarmupCosine(optim, optim_cfg.MAX_EPOCH, warm, getattr(optim_cfg, "WARMUP_CONS_LR", 1e-5))


# ---------------------------------------------------------------------------------------------------- data parallel
class DataParallelGroup:
    """One process per GPU.  Images shard by batch; the frozen CLIP weights are resident on every rank; the only
    per-step exchange is a SUM all-reduce of the flat prompt-gradient buffer (each rank's cross-entropy already divides
    by the GLOBAL batch, so the sum equals the single-GPU gradient — SURVEY.md §8e)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.enabled = dist.is_available() and dist.is_initialized()
        self.group = group
        self.world = dist.get_world_size(group) if self.enabled else 1
        self.rank = dist.get_rank(group) if self.enabled else 0

    @staticmethod
    def from_env(backend: Optional[str] = None) -> "DataParallelGroup":
        """Initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / MASTER_*)."""
        import torch.distributed as dist
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world > 1 and not dist.is_initialized():
            backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            if backend == "nccl":
                torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
            dist.init_process_group(backend=backend, rank=int(os.environ["RANK"]), world_size=world)
        return DataParallelGroup()

    def all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        if self.enabled and self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def all_reduce_max(self, t: torch.Tensor) -> torch.Tensor:
        if self.enabled and self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return t

    def barrier(self):
        if self.enabled and self.world > 1:
            self.dist.barrier(group=self.group)

    def shard(self, n: int, rank: Optional[int] = None) -> Tuple[int, int]:
        """[start, end) of a rank's slice of n items (balanced, contiguous); this rank's by default."""
        rank = self.rank if rank is None else rank
        base, rem = divmod(n, self.world)
        start = rank * base + min(rank, rem)
        return start, start + base + (1 if rank < rem else 0)

    def all_gather_into(self, out: torch.Tensor, send: torch.Tensor) -> torch.Tensor:
        """out[r] = rank r's `send` (out: [world, *send.shape]) — the text-feature exchange of the class-sharded tower."""
        if self.enabled and self.world > 1:
            self.dist.all_gather_into_tensor(out.view(-1), send.view(-1), group=self.group)
        else:
            out.view(-1).copy_(send.view(-1))
        return out

    def reduce_scatter_sum(self, out: torch.Tensor, send: torch.Tensor) -> torch.Tensor:
        """out = sum over ranks of their send[this rank] (send: [world, *out.shape])."""
        if self.enabled and self.world > 1:
            self.dist.reduce_scatter_tensor(out.view(-1), send.view(-1), op=self.dist.ReduceOp.SUM, group=self.group)
        else:
            out.view(-1).copy_(send.view(-1))
        return out

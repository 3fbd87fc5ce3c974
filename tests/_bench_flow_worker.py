"""Worker of tests/test_cpu_bench_flow.py: runs bench.py's GPU arm — its real control flow: warm-up, agreed thermal settle,
the three timed regions with their barriers, FLOP accounting, the JSON line — in a container WITHOUT a GPU, one process per
rank over gloo.  Everything that needs the device is replaced HERE, in the test process only (bench.py has no dry-run
switch): `torch.cuda` by host-clock stand-ins, the trainer by a stub whose step issues the same collective a training
step does (one SUM all-reduce of the prompt-gradient buffer) and takes a rank-dependent time, so that ranks which did not
agree on their step counts would deadlock here exactly as they would on NCCL."""
import os
import sys
import time
import types

import torch

REPO = os.environ["REPO"]
sys.path.insert(0, REPO)
import bench  # noqa: E402
from mvlpt_b200 import _lib  # noqa: E402
from mvlpt_b200.trainers import mvlpt as trainer_mod  # noqa: E402
from mvlpt_b200.trainers import runtime  # noqa: E402


class HostEvent:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3

    def synchronize(self):
        pass


class TorchStandIn(types.ModuleType):
    """`torch` as bench.py sees it: the real module, except that devices are the CPU and torch.cuda is a host stand-in."""

    def __init__(self):
        super().__init__("torch")
        self.cuda = types.SimpleNamespace(set_device=lambda d: None, synchronize=lambda *a: None, Event=HostEvent,
                                          empty_cache=lambda: None, is_available=lambda: False)

    def device(self, *a):
        return torch.device("cpu")

    def __getattr__(self, k):
        return getattr(torch, k)


class SkewedClock(types.ModuleType):
    """`time` as bench.py sees it on this rank: the host clocks of two ranks never agree — rank r's runs 7 % x r fast."""

    def __init__(self, rank):
        super().__init__("time")
        self.rate = 1.0 + 0.07 * rank

    def perf_counter(self):
        return time.perf_counter() * self.rate

    def __getattr__(self, k):
        return getattr(time, k)


class StubTrainer:
    steps = 0

    def __init__(self, cfg, dm=None, clip_state_dict=None, device=None, tokenized_prompts=None, name_lens=None, dp=None):
        self.dp, self.num_batches = dp, 1
        self.model = types.SimpleNamespace(
            prompt_learner=types.SimpleNamespace(load_state_dict=lambda *a, **k: None, kernel_len=25),
            shard_text=True, hold_text_features=lambda on: None)
        self.grad = torch.zeros(8192)

    def _normalize_u8(self, img):
        return img.float()

    def set_model_mode(self, mode):
        pass

    def stage_batch(self, batch):
        return batch

    def parse_batch_test(self, batch):
        return batch["img"], batch["label"], batch["domain"]

    def model_inference(self, image, task=None):
        StubTrainer.steps += 1
        return torch.zeros(image.shape[0], 16)

    def forward_backward(self, batch):
        StubTrainer.steps += 1
        time.sleep(0.002 * (1 + self.dp.rank))  # ranks run at different speeds: their own clocks disagree
        self.grad.fill_(1.0)
        self.dp.all_reduce_sum(self.grad)       # the collective of a training step
        assert float(self.grad[0]) == self.dp.world
        return trainer_mod.LossSummary(torch.tensor([1.0, 50.0]), HostEvent(), {})


def main():
    torch.Tensor.pin_memory = lambda self, *a, **k: self   # no accelerator here
    bench.torch = TorchStandIn()
    bench.time = SkewedClock(int(os.environ.get("RANK", "0")))
    trainer_mod.MVLPT = StubTrainer
    _lib.check = lambda rc, what="": None
    _lib.lib = lambda: types.SimpleNamespace(mvlpt_check_device=lambda i: 0)
    _lib.launch_count = lambda: StubTrainer.steps
    real_from_env = runtime.DataParallelGroup.from_env
    runtime.DataParallelGroup.from_env = classmethod(lambda cls, backend="nccl": real_from_env("gloo"))
    # a small stand-in for the CLIP weights: the stub trainer never reads them
    bench.synth.synth_clip_state_dict = lambda arch, seed=0: {}
    sys.argv = ["bench.py"] + sys.argv[1:]
    bench.main()
    print(f"rank {os.environ.get('RANK', '0')} steps {StubTrainer.steps}", file=sys.stderr)


if __name__ == "__main__":
    main()

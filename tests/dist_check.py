"""Multi-GPU check run under torchrun on a GPU box (scripts/gpu_multi.sh): the sharded enhancement of N clips over
the ranks + one NCCL all-gather must equal the single-GPU enhancement of the same clips."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fullsubnet-plus_b200")]
from fsnplus_b200 import inference as inf  # noqa: E402
from fsnplus_b200.model import FullSubNet_Plus  # noqa: E402
from fsnplus_b200.synth import synth_clips  # noqa: E402
import bench  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = FullSubNet_Plus(**bench.default_cfg()).to(dev).eval()
    clips = synth_clips(7, 48000, 16000, seed=4242).to(dev)            # ragged: 7 clips over 2 ranks
    out = inf.enhance_sharded(model, clips)
    ref = inf.enhance_batch(model, clips)
    err = (out - ref).norm() / ref.norm()
    ok = out.shape == ref.shape and err.item() < 1e-5
    # the pipelined path of bench.py: every rank pushes its own batches, the all-gather of batch i runs on the side stream under the
    # forward of batch i+1; result k must be the concatenation of all ranks' enhanced batch k
    mine = [synth_clips(3, 48000, 16000, seed=100 * k + rank).to(dev) for k in range(3)]
    pipe = inf.EnhancePipeline(model, 48000)
    for b in mine:
        pipe.push(inf.stft(b))
    got = [r.clone() for r in pipe.flush()]
    for k in range(3):
        want = torch.cat([inf.enhance_batch(model, synth_clips(3, 48000, 16000, seed=100 * k + r).to(dev)) for r in range(world)], 0)
        e2 = ((got[k] - want).norm() / want.norm()).item()
        ok = ok and got[k].shape == want.shape and e2 < 1e-5
        err = torch.maximum(err, torch.tensor(e2, device=dev))
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"dist_check world={world} shape={tuple(out.shape)} rel_err={err.item():.2e} ok={bool(flag.item())}")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()

"""N>1 host logic on CPU: frame sharding + the confusion-matrix all-reduce (gloo, world_size 2).
The device kernels are not involved: each rank fills its ConfusionMeter from the oracle's fast_hist of its
shard, then the same ConfusionMeter.all_reduce() the GPU path uses must reproduce the global matrix."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import drnb200
from oracle import drn_oracle


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(0)
    pred = rng.randint(0, 19, size=(n_frames, 16, 32))
    label = rng.randint(0, 21, size=(n_frames, 16, 32))
    label[label >= 19] = 255
    lo, hi = drnb200.shard_frames(n_frames, rank, world)
    meter = drnb200.ConfusionMeter(19, torch.device("cpu"))
    if hi > lo:
        meter.hist += torch.from_numpy(drn_oracle.fast_hist(pred[lo:hi].flatten(), label[lo:hi].flatten(), 19))
    meter.all_reduce()
    full = drn_oracle.fast_hist(pred.flatten(), label.flatten(), 19)
    ok = np.array_equal(meter.hist.numpy(), full) and meter.miou() == drn_oracle.miou(full)
    out[rank] = int(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_confusion_allreduce_world2():
    for n_frames in (7, 1):          # ragged split, and a rank with no frames at all
        with mp.Manager() as mgr:
            out = mgr.dict()
            mp.spawn(_worker, args=(2, _free_port(), n_frames, out), nprocs=2, join=True)
            assert dict(out) == {0: 1, 1: 1}


def test_all_reduce_is_identity_without_process_group():
    meter = drnb200.ConfusionMeter(3, torch.device("cpu"))
    meter.hist += 2
    assert int(meter.all_reduce().sum()) == 18

"""World-size-2 (gloo, CPU) check of the N>1 host logic of bench.py: batch shards are independent,
per-rank partial means average to the global-batch loss, and the max-over-ranks reduction works.
The loss arithmetic here is the oracle (CPU); the GPU kernels need no collective."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from monodepth2_b200.synthetic import make_batch
    from oracle import view_synthesis as O
    B, H, W, fids = 4, 32, 64, [0, -1, 1]
    inputs, outputs, pose, noise = make_batch(B, H, W, fids, 4, seed=11, kind="structured")
    per = B // world
    sl = slice(rank * per, (rank + 1) * per)
    cfg = O.OracleConfig(height=H, width=W, frame_ids=tuple(fids))
    ins = {k: v[sl] for k, v in inputs.items()}
    outs = {k: v[sl].clone() for k, v in outputs.items()}
    loss = O.view_synthesis_loss(ins, outs, cfg, [n[sl] for n in noise])["loss"].detach().double()
    t = torch.tensor([float(rank + 1) * 1.5])          # stand-in for the per-rank device time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(loss, op=dist.ReduceOp.SUM)
    if rank == 0:
        full = O.view_synthesis_loss(dict(inputs), {k: v.clone() for k, v in outputs.items()}, cfg, noise)["loss"]
        out.put((float(loss / world), float(full), float(t)))
    dist.barrier()
    dist.destroy_process_group()


def test_batch_shards_average_to_global_loss():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    sharded, full, tmax = q.get()
    # the photometric term is a per-rank mean; the smoothness mean-normalisation is per-sample: shards are exact
    assert abs(sharded - full) <= 1e-6 * abs(full)
    assert tmax == 3.0


def test_reference_arm_exits_quietly_on_nonzero_rank():
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""

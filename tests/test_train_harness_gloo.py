"""World-size-2 (gloo, CPU) check of the N>1 training-step harness of bench.py --mode train
(benchmarks/train_step.py): DDP over the container of the four stand-in networks, the pose encoder
run twice per forward, one backward, Adam.  The loss on this CPU leg is the oracle; on the GPU box
the same harness calls the fused loss.  Checks that every parameter receives a gradient, that the
ranks hold identical weights after two steps on different data, and that the weights moved."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from torch.nn.parallel import DistributedDataParallel as DDP
    from benchmarks.train_step import Nets, TrainStep
    from monodepth2_b200.synthetic import make_batch
    from oracle import view_synthesis as O
    B, H, W, fids = 2, 64, 96, [0, -1, 1, "s"]
    torch.manual_seed(7)
    nets = Nets(fids, 18)
    w0 = torch.cat([p.detach().flatten() for p in nets.parameters()]).clone()
    model = DDP(nets)
    cfg = O.OracleConfig(height=H, width=W, frame_ids=tuple(fids))
    step = TrainStep(model, fids, lambda i, o: O.view_synthesis_loss(i, o, cfg, None),
                     O.transformation_from_parameters)
    losses = []
    for i in range(2):
        inputs, _, _, _ = make_batch(B, H, W, fids, 4, seed=100 * rank + i, kind="structured")
        losses.append(float(step(inputs)))
    missing = [n for n, p in nets.named_parameters() if p.grad is None]
    w1 = torch.cat([p.detach().flatten() for p in nets.parameters()])
    gathered = [torch.empty_like(w1) for _ in range(world)]
    dist.all_gather(gathered, w1)
    if rank == 0:
        out.put((missing, float((gathered[0] - gathered[1]).abs().max()), float((w1 - w0).abs().max()), losses))
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_training_step_keeps_ranks_in_sync():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 29700 + (os.getpid() % 250)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    missing, rank_diff, moved, losses = q.get()
    assert missing == []                      # no unused parameters (DDP find_unused_parameters=False holds)
    assert rank_diff == 0.0                   # all-reduced gradients -> identical weights on both ranks
    assert moved > 0.0
    assert all(l == l and l > 0 for l in losses)

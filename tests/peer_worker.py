"""One rank of the peer-memory exchange test (tests/test_gpu_peer.py launches WORLD_SIZE of these; gloo carries the IPC
handles and the cross-checks, so two ranks may share one GPU when the box has a single device)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "supervised-depth-estimation-from-polarized-images_b200"))
from polcue import dist as D, ops, synth  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", rank % torch.cuda.device_count())
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    peer = D.PeerExchange(dev)
    assert (peer.rank, peer.world) == (rank, world)

    def rank_order_sum(t):
        parts = [torch.empty_like(t.cpu()) for _ in range(world)]
        dist.all_gather(parts, t.cpu())
        total = parts[0].clone()
        for p in parts[1:]:
            total += p                                     # the order the kernel adds in
        return total

    # stand-alone exchange: many back-to-back calls without host synchronisation (parity double buffering), every length class
    gen = torch.Generator().manual_seed(100 + rank)
    calls = 0
    for n in (1, 7, 78, 113, 128):
        vals = [torch.randn(n, dtype=torch.float64, generator=gen).to(dev) * 10.0 ** (i % 5) for i in range(12)]
        outs = [peer.all_reduce(v) for v in vals]
        calls += len(vals)
        torch.cuda.synchronize()
        for v, o in zip(vals, outs):
            assert torch.equal(o.cpu(), rank_order_sum(v)), (rank, n)
    # NaN and inf travel like any other value
    v = torch.tensor([float("nan") if rank == 0 else 1.0, float("inf"), 3.0], dtype=torch.float64, device=dev)
    o = peer.all_reduce(v).cpu()
    calls += 1
    assert np.isnan(o[0].item()) and np.isinf(o[1].item()) and o[2].item() == 3.0 * world

    # the evaluation pass with the exchange fused into its last kernel == the plain pass + a rank-ordered sum
    images = 6
    lo, hi = D.shard_range(images, rank, world)
    gt, pred, inst, k = (torch.from_numpy(a).to(dev) for a in synth.gen_depth_batch(lo, hi - lo))
    groups = [None] + list(synth.MATERIAL_LEVELS)
    plain = ops.eval_pass(gt, pred, inst, k, 0.1, 2.0, groups)
    fused = ops.eval_pass(gt, pred, inst, k, 0.1, 2.0, groups, peer=peer)
    calls += 1
    torch.cuda.synchronize()
    for key in ("normals", "sums", "metrics", "mean_acc"):
        assert torch.equal(plain[key], fused[key]), key
    want = rank_order_sum(plain["mean_acc"])
    assert torch.equal(fused["mean_acc_all"].cpu(), want), (fused["mean_acc_all"], want)
    assert want[0].item() == images

    # replayed from a CUDA graph: the call counter lives on the device, so every replay is a new exchange
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            fused = ops.eval_pass(gt, pred, inst, k, 0.1, 2.0, groups, out=fused, peer=peer)
    torch.cuda.current_stream(dev).wait_stream(side)
    for _ in range(5):
        fused["mean_acc_all"].zero_()
        g.replay()
        calls += 1
        torch.cuda.synchronize()
        assert torch.equal(fused["mean_acc_all"].cpu(), want)
    made, failed = peer.status()
    assert (made, failed) == (calls, 0), (made, failed, calls)
    peer.close()
    dist.barrier()
    print(f"PEER_OK rank {rank} of {world} on {dev}, {calls} exchanges", flush=True)


if __name__ == "__main__":
    main()

"""The one collective of the path over NVLink peer memory (csrc/peer.cuh, polcue_peer_* / polcue_eval_pass_peer_f32)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "supervised-depth-estimation-from-polarized-images_b200"))

pytestmark = pytest.mark.gpu


def test_single_rank_exchange_is_the_identity_counts_its_calls_and_rejects_bad_arguments():
    from polcue import dist as D, ops, synth
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    peer = D.PeerExchange(dev)
    assert (peer.rank, peer.world) == (0, 1)
    v = torch.randn(78, dtype=torch.float64, device=dev)
    assert torch.equal(peer.all_reduce(v), v)
    out = torch.empty_like(v)
    assert peer.all_reduce(v, out=out) is out and torch.equal(out, v)
    for bad in (v.float(), torch.randn(129, dtype=torch.float64, device=dev), v.cpu(), v[::2]):
        with pytest.raises(ValueError):
            peer.all_reduce(bad)
    gt, pred, inst, k = (torch.from_numpy(a).to(dev) for a in synth.gen_depth_batch(0, 5))
    groups = [None] + list(synth.MATERIAL_LEVELS)
    plain = ops.eval_pass(gt, pred, inst, k, 0.1, 2.0, groups)
    fused = ops.eval_pass(gt, pred, inst, k, 0.1, 2.0, groups, peer=peer)
    for key in ("normals", "sums", "metrics", "mean_acc"):
        assert torch.equal(plain[key], fused[key]), key
    assert torch.equal(fused["mean_acc_all"], plain["mean_acc"])
    # the prepared launcher binds arguments and buffers once; inputs are captured by reference
    prepared = ops.EvalPass(gt, pred, inst, k, 0.1, 2.0, groups, peer=peer)
    got = prepared.run()
    for key in ("normals", "sums", "metrics", "mean_acc"):
        assert torch.equal(plain[key], got[key]), key
    assert torch.equal(got["mean_acc_all"], plain["mean_acc"])
    pred.mul_(1.01)
    again = ops.eval_pass(gt, pred, inst, k, 0.1, 2.0, groups)
    got = prepared.run()
    assert torch.equal(got["mean_acc_all"], again["mean_acc"]) and not torch.equal(again["mean_acc"], plain["mean_acc"])
    assert peer.status() == (5, 0)
    peer.close()
    peer.close()                                   # idempotent


@pytest.mark.parametrize("world", [2, 4])
def test_exchange_between_processes_is_the_rank_ordered_sum(world):
    """WORLD processes (one GPU each when the box has them, else sharing cuda:0 -- IPC works across processes either way):
    stand-alone exchanges of every length class back to back, the fused evaluation pass, CUDA-graph replays."""
    if world > 2 and torch.cuda.device_count() < world:
        pytest.skip("four ranks time-slicing one GPU only repeat the two-rank case slowly")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "peer_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=600)[0])
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"PEER_OK rank {rank} of {world}" in o, o[-3000:]

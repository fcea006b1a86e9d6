"""world_size-2 data-parallel host logic on CPU (gloo): batch sharding + global-count loss
normalisation + bucketed flat all-reduce must reproduce the single-process gradient of the
concatenated batch.  The per-rank compute stand-in is the oracle."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodal_image_transformer_b200 import dp
from oracle import decoder_oracle as O
from tests.helpers import CFGS, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, _, w = dp.DataParallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    torch.set_num_threads(2)
    c = dict(CFGS["nano"], B=5)
    p = O.init_params(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], seed=11)
    tok, tgt, mem, _ = synth(c, seed=5)
    b0, b1 = dp.shard_range(c["B"], rank, world)
    loss, grads = O.loss_and_grads(p, tok[b0:b1], tgt[b0:b1], mem[b0:b1], None, c["H"])
    n_local = (tgt[b0:b1] != 0).sum().float().view(1)
    n_global = n_local.clone()
    dist.all_reduce(n_global)
    # flat arena in a fixed name order, gradient of the GLOBAL mean = local mean-grad * n_local/n_global
    names = sorted(grads)
    sizes = [grads[k].numel() for k in names]
    flat = torch.cat([grads[k].flatten() for k in names]) * (n_local / n_global)
    # buckets: segments in decreasing address order (the order backward finishes them)
    offs, segs, o = [], [], 0
    for n in sizes:
        offs.append(o)
        o += n
    for off, n in reversed(list(zip(offs, sizes))):
        segs.append((off, n))
    buckets = dp.plan_buckets(segs, cap_elems=20000)
    assert sum(cn for _, cn in buckets) == flat.numel() and len(buckets) > 1
    dp.allreduce_flat(flat, buckets)
    loss_sum = (loss * n_local).clone()
    dist.all_reduce(loss_sum)
    if rank == 0:
        torch.save({"flat": flat, "names": names, "sizes": sizes, "loss": (loss_sum / n_global).item()},
                   os.path.join(out_dir, "dp.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_equals_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = torch.load(os.path.join(str(tmp_path), "dp.pt"), weights_only=True)
    c = dict(CFGS["nano"], B=5)
    p = O.init_params(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], seed=11)
    tok, tgt, mem, _ = synth(c, seed=5)
    loss, grads = O.loss_and_grads(p, tok, tgt, mem, None, c["H"])
    ref = torch.cat([grads[k].flatten() for k in got["names"]])
    assert abs(got["loss"] - loss.item()) < 1e-5 * loss.item()
    assert ((got["flat"] - ref).norm() / ref.norm()).item() < 1e-5

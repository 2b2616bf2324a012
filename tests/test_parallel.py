"""N>1 host logic on CPU: two gloo ranks shard a batch through the product's ShardedAero (kernels emulated by
tests/cpu_emu.py) and must reproduce the single-process result; plus the shard arithmetic."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aero_b200.parallel import shard_range

HERE = os.path.dirname(os.path.abspath(__file__))


def test_shard_range_is_a_partition():
    for n in (1, 5, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [h - l for l, h in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cpu_emu import emulated
    from util import SEED, trained_like_, white_noise
    from aero_b200 import Aero, aero_kwargs
    from aero_b200.parallel import ShardedAero, reduce_max
    torch.manual_seed(SEED)
    m = Aero(**aero_kwargs("aero_4-16_512_256")).eval()
    m.load_state_dict(trained_like_(m.state_dict()))
    emulated(m)
    mix = white_noise((3, 1, 4000))
    sh = ShardedAero(m)
    full = sh.forward(mix, gather=True)
    local = sh.forward(mix)
    t = reduce_max(1.0 + rank)
    if rank == 0:
        q.put((full, local.shape[0], t))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_forward_matches_single_process():
    sys.path.insert(0, HERE)
    from cpu_emu import emulated
    from util import SEED, rel_l2, trained_like_, white_noise
    from aero_b200 import Aero, aero_kwargs
    torch.manual_seed(SEED)
    m = Aero(**aero_kwargs("aero_4-16_512_256")).eval()
    m.load_state_dict(trained_like_(m.state_dict()))
    emulated(m)
    ref = m(white_noise((3, 1, 4000)))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, n_local, t = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert n_local == 2                      # rank 0 gets the extra clip of 3
    assert t == pytest.approx(2.0)           # max over ranks
    assert full.shape == ref.shape and rel_l2(full, ref) < 1e-6


def test_trainer_flat_buffer_order_and_pieces():
    """Host logic of aero_b200.trainer (no kernels): gradients are laid out in the order the backward pass finishes them and the
    all-reduce pieces end at layer boundaries and tile the buffer."""
    from aero_b200 import Aero, aero_kwargs
    from aero_b200.trainer import GeneratorTrainer, _backward_order
    torch.manual_seed(0)
    m = Aero(**aero_kwargs("aero_4-16_512_256"))
    order = _backward_order(m)
    assert sorted(order) == sorted(n for n, _ in m.named_parameters())
    heads = [".".join(n.split(".")[:2]) for n in order]
    first = {h: heads.index(h) for h in dict.fromkeys(heads)}
    want = [f"decoder.{j}" for j in (3, 2, 1, 0)] + [f"encoder.{i}" for i in (3, 2, 1, 0)]
    assert [h for h in first if h.startswith(("decoder", "encoder"))] == want
    tr = GeneratorTrainer.__new__(GeneratorTrainer)
    tr.order = order
    tr.flat = torch.zeros(sum(p.numel() for p in m.parameters()))
    tr.offsets, off = {}, 0
    params = dict(m.named_parameters())
    for n in order:
        tr.offsets[n] = (off, off + params[n].numel())
        off += params[n].numel()
    pieces = tr._make_pieces(4)
    assert pieces[0][0] == 0 and pieces[-1][1] == tr.flat.numel() and all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
    ends = {tr.offsets[n][1] for n in order}
    assert all(hi in ends for _, hi in pieces) and 2 <= len(pieces) <= 8

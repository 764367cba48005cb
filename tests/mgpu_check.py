"""Run under torchrun (one rank per GPU): N-GPU results vs the 1-GPU / oracle results.
  * heat: slab-decomposed run with NCCL halo exchange must be BIT-IDENTICAL to the undivided
    grid (slabbing changes no cell's arithmetic) -- checked against the oracle on a small grid
    and against a single-GPU run of the same library on a larger one;
  * reductions: allreduce of per-GPU partials (sum on exactly-summable data: exact; max: exact;
    argmax: first extremum across shards).
Prints MGPU_OK on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ph_core_b200 as ph
from ph_core_b200 import DeviceNArray as D, sharding as S, heat
from oracle import ph_oracle as O


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ph.init(local)
    world, rank = S.comm_init(dist)
    rs = np.random.RandomState(7)

    # ---- heat, small grid vs oracle (uneven split); one ghost plane (a step per exchange) and two
    #      (two time steps per pass and per exchange), odd and even step counts
    field = (rs.rand(4 * world + 3, 20, 36) * 100).astype(np.float32)
    for ghost, steps in ((1, 7), (2, 7), (2, 6)):
        want = field.copy()
        for _ in range(steps):
            want = O.heat_step_nd(want, np.float32(0.1))
        lay = S.slab_layout(field.shape[0], world, rank, ghost)
        loc = S.slab_from_global(field, world, rank, ghost)
        a, b = D.from_host(loc), D.from_host(loc)
        fin = S.heat_run_sharded(a, b, 0.1, steps, ghost).to_host()[ghost:-ghost]
        assert fin.tobytes() == want[lay["start"]:lay["stop"]].tobytes(), \
            f"rank {rank}: small heat slab (ghost {ghost}, {steps} steps) differs from oracle"

    # ---- heat, larger grid vs a single-GPU run of the same library (rank 0 computes it)
    big = (rs.rand(16 * world, 96, 256) * 100).astype(np.float32)
    for ghost, steps in ((1, 9), (2, 9), (2, 12)):
        lay = S.slab_layout(big.shape[0], world, rank, ghost)
        loc = S.slab_from_global(big, world, rank, ghost)
        a, b = D.from_host(loc), D.from_host(loc)
        fin = S.heat_run_sharded(a, b, 0.1, steps, ghost).to_host()[ghost:-ghost]
        whole = heat.simulate(D.from_host(big), 0.1, steps).to_host()
        assert fin.tobytes() == whole[lay["start"]:lay["stop"]].tobytes(), \
            f"rank {rank}: big heat slab (ghost {ghost}, {steps} steps) differs from 1-GPU run"

    # ---- 2-D grid (BASELINE configs[0] generalised), two ghost rows, two steps per pass
    flat2 = (rs.rand(64 * world + 5, 520) * 100).astype(np.float32)
    for ghost, steps in ((2, 8), (2, 5), (1, 3)):
        want = flat2.copy()
        for _ in range(steps):
            want = O.heat_step_nd(want, np.float32(0.1))
        lay = S.slab_layout(flat2.shape[0], world, rank, ghost)
        loc = S.slab_from_global(flat2, world, rank, ghost)
        a, b = D.from_host(loc), D.from_host(loc)
        fin = S.heat_run_sharded(a, b, 0.1, steps, ghost).to_host()[ghost:-ghost]
        assert fin.tobytes() == want[lay["start"]:lay["stop"]].tobytes(), \
            f"rank {rank}: 2-D heat slab (ghost {ghost}, {steps} steps) differs from oracle"

    # ---- sharded reductions
    data = rs.randint(-8, 9, size=(8 * world + 1, 50, 30)).astype(np.float32)
    data[3, 2, 1] = data[-1, 4, 4] = 99.0                      # tie across shards
    r0, r1 = S.shard_range(data.shape[0], world, rank)
    x = D.from_host(data[r0:r1])
    off = r0 * 50 * 30
    assert S.reduce_full_sharded(x, "sum") == np.float32(data.sum(dtype=np.float64))
    assert S.reduce_full_sharded(x, "max") == data.max() and S.reduce_full_sharded(x, "min") == data.min()
    v, i = S.reduce_full_sharded(x, "argmax", off)
    assert (v, i) == (np.float32(99.0), 3 * 1500 + 2 * 30 + 1), (v, i)
    # axis-0 reduce: allreduce of the [outer*inner] partial
    part = x.sum(axis=0)
    ph.check(ph.load().ph_allreduce(ph.K["PH_SUM"], ph.K["PH_F32"], part.ptr, part.size))
    assert part.to_host().tobytes() == data.sum(axis=0, dtype=np.float64).astype(np.float32).tobytes()
    # ---- f-3: ShardedNArray behaves like the undivided array
    g = rs.randint(-8, 9, size=(6 * world + 2, 12, 10)).astype(np.float32)
    h = rs.randint(-8, 9, size=(6 * world + 2, 12, 10)).astype(np.float32)
    sg, sh = S.ShardedNArray.from_global(g), S.ShardedNArray.from_global(h)
    assert ((sg * sh + sg) - 2.0).to_global().tobytes() == ((g * h + g) - np.float32(2.0)).tobytes()
    assert (sg > sh).to_global().tobytes() == (g > h).tobytes()
    assert sg.sum() == np.float32(g.sum(dtype=np.float64)) and sg.max() == g.max()
    flat = int(np.argmax(g.reshape(-1)))
    assert sg.argmax() == (g.max(), list(np.unravel_index(flat, g.shape)))
    assert sg.sum(axis=0).to_host().tobytes() == g.sum(axis=0, dtype=np.float64).astype(np.float32).tobytes()
    assert sg.max(axis=2).to_global().tobytes() == g.max(axis=2).tobytes()
    assert sg[ph.ALL, ph.rng(1, 9, 2), 3].to_global().tobytes() == np.ascontiguousarray(g[:, 1:10:2, 3]).tobytes()
    # transposes across shards (ph_alltoallv): default pattern = reversed axes, then two others
    assert sg.permute().to_global().tobytes() == np.ascontiguousarray(g.transpose(2, 1, 0)).tobytes()
    assert sg.permute(1, 0, 2).to_global().tobytes() == np.ascontiguousarray(g.transpose(1, 0, 2)).tobytes()
    assert sg.permute(0, 2, 1).to_global().tobytes() == np.ascontiguousarray(g.transpose(0, 2, 1)).tobytes()
    m2 = rs.rand(1000 * world + 3, 517).astype(np.float64)
    t2 = S.ShardedNArray.from_global(m2).permute()
    assert t2.shape == [517, 1000 * world + 3] and t2.to_global().tobytes() == np.ascontiguousarray(m2.T).tobytes()
    assert t2.permute().to_global().tobytes() == m2.tobytes()              # transposing twice is the identity
    sg.set_mask(sg > sh, 0.0)
    assert sg.to_global().tobytes() == np.where(g > h, np.float32(0), g).tobytes()
    dist.barrier()
    if rank == 0:
        print(f"MGPU_OK world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""world_size-2 gloo tests of the multi-GPU host logic (no GPU): env sharding keeps per-env streams, and the
sum all-reduce of per-shard gradients (mean loss divided by the GLOBAL batch) is the global gradient."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_envs_partition():
    from dqnflappybird_b200.dist import owner_of_env, shard_envs
    for total, world in ((4096, 1), (65536, 8), (1048576, 8), (10, 4), (7, 2)):
        seen = []
        for r in range(world):
            first, n = shard_envs(total, r, world)
            seen += list(range(first, first + n))
            if n:
                assert owner_of_env(first, total, world) == r and owner_of_env(first + n - 1, total, world) == r
        assert seen == list(range(total))
    assert shard_envs(1048576, 3, 8) == (3 * 131072, 131072)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from dqnflappybird_b200 import dist as fd
    from dqnflappybird_b200.assets import load_blob
    from dqnflappybird_b200 import _lib
    from oracle import qnet_oracle as qo
    import ctypes as C
    r, w, _ = fd.init("gloo")
    assert (r, w) == (rank, world)
    # (1) env sharding: a shard's envs are the global envs with the same ids (Philox stream = (seed, env id))
    L = _lib.lib()
    blob = load_blob()
    assert L.fb_debug_assets_load_host(blob, len(blob)) == 0
    total = 6
    first, n = fd.shard_envs(total, rank, world)
    states = []
    for e in range(first, first + n):
        st = np.zeros(16, np.int32)
        L.fb_debug_host_reset(st.ctypes.data, None, 0, 42, e)
        rr, tt, ss = C.c_float(), C.c_uint8(), C.c_int32()
        for k in range(120):
            L.fb_debug_host_step(st.ctypes.data, int(k % 7 == 0), None, 0, 42, e, C.byref(rr), C.byref(tt), C.byref(ss))
        states.append(st)
    gathered = [None] * world
    dist.all_gather_object(gathered, (first, np.stack(states)))
    # (2) gradients: shard of B/world samples, loss divided by the global batch, sum all-reduce
    B = 8
    rng = np.random.default_rng(0)                      # same data on every rank
    x = (rng.random((B, 5, 80, 80)) < 0.2).astype(np.uint8) * 255
    p = qo.init_params(64, False, 1) * np.float32(3); t = qo.init_params(64, False, 2) * np.float32(3)
    a = rng.integers(0, 2, B).astype(np.uint8); rew = rng.choice(np.array([0.1, 3, -3], np.float32), B); term = (rew == -3).astype(np.uint8)
    lb = fd.local_batch(B, world)
    sl = slice(rank * lb, (rank + 1) * lb)
    loss, g, *_ = qo.loss_and_grads(1, p, t, x[sl, 0:4], x[sl, 1:5], a[sl], rew[sl], term[sl], None, 0.99, False, B, 64, False)
    gt = torch.from_numpy(g)
    fd.allreduce_gradients(gt)
    lt = torch.tensor([loss], dtype=torch.float64)
    dist.all_reduce(lt)
    if rank == 0:
        loss_full, g_full, *_ = qo.loss_and_grads(1, p, t, x[:, 0:4], x[:, 1:5], a, rew, term, None, 0.99, False, None, 64, False)
        q.put((gathered, float(lt.item()), loss_full, float(np.abs(gt.numpy() - g_full).max()), float(np.abs(g_full).max())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_sharding_and_gradient_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, loss_sum, loss_full, gerr, gmax = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert abs(loss_sum - loss_full) <= 1e-12 * abs(loss_full)
    assert gerr <= 1e-12 * gmax
    # the two shards together are exactly the 6 global envs run in one process
    sys.path.insert(0, ROOT)
    from oracle import flappy_oracle as fo
    env = fo.OracleEnvs(6, seed=42)
    for k in range(120):
        env.step(np.full(6, int(k % 7 == 0), np.uint8), want_obs=False)
    want = env.export_state()
    got = np.concatenate([s for _, s in sorted(gathered, key=lambda fs: fs[0])])
    np.testing.assert_array_equal(got, want)

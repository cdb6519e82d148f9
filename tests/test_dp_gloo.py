"""world_size-2 `gloo` tests (CPU) of the data-parallel host logic (spfsplatv2_b200/dp.py): view sharding and the
one gradient all-reduce per step.  The renderer stand-in is the CPU oracle -- the test exercises the plumbing, the
CUDA path is covered by the `-m gpu` tests and bench.py --gpus N."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spfsplatv2_b200.dp import GradAllReduce, render_views_sharded, shard_indices


def test_shard_indices_partition_ragged_and_empty():
    for n, world in [(16, 2), (7, 4), (3, 8), (0, 2), (8, 8)]:
        parts = [shard_indices(n, r, world) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert shard_indices(3, 5, 8) == []
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _scene_and_views(n_views):
    from spfsplatv2_b200.synthetic import make_scene
    return make_scene(seed=41, v_cxt=1, h=32, w=32, grid=(10, 10), regime="trained", n_target=n_views)


def _render_fns(sc, leaves):
    from oracle import raster_oracle as O
    from spfsplatv2_b200.camera import camera_setup
    h, w = sc.image_shape
    view, proj, tanfov, scale = camera_setup(sc.extrinsics[0], sc.intrinsics[0], sc.near[0], sc.far[0], True)
    wts = torch.randn(sc.extrinsics.shape[1], 3, h, w, generator=torch.Generator().manual_seed(5))

    def render_view(i):
        vw = O.View(h, w, float(tanfov[i, 0]), float(tanfov[i, 1]), torch.zeros(3), view[i].contiguous(), proj[i].contiguous(), 4, 1.0)
        return O.render(leaves["means"] * scale[i], leaves["scales"] * scale[i], leaves["rotations"], leaves["opacities"],
                        leaves["harmonics"].permute(0, 2, 1).contiguous(), None, vw)["color"]

    def loss_of_view(i, color):
        return (color * wts[i]).sum()
    return render_view, loss_of_view


def _leaves(sc):
    return {k: getattr(sc, k)[0].clone().requires_grad_() for k in ("means", "scales", "rotations", "opacities", "harmonics")}


def _worker(rank, world, port, n_views, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        sc = _scene_and_views(n_views)
        leaves = _leaves(sc)
        rv, lv = _render_fns(sc, leaves)
        red = GradAllReduce(torch.device("cpu"))
        grads = render_views_sharded(rv, lv, n_views, leaves, red)
        torch.save({k: v for k, v in grads.items()}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_views", [3, 1])     # ragged (2+1) and "rank 1 has nothing to render"
def test_sharded_views_allreduce_equals_single_process(tmp_path, n_views):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_views, str(tmp_path)), nprocs=world, join=True)
    sc = _scene_and_views(n_views)
    leaves = _leaves(sc)
    rv, lv = _render_fns(sc, leaves)
    total = sum(lv(i, rv(i)) for i in range(n_views))
    total.backward()
    got = [torch.load(os.path.join(str(tmp_path), f"rank{r}.pt")) for r in range(world)]
    for k, t in leaves.items():
        for r in range(world):
            assert torch.allclose(got[r][k], t.grad, rtol=1e-5, atol=1e-6 * float(t.grad.abs().max())), (k, r)
        assert torch.equal(got[0][k], got[1][k])          # every rank ends with the same reduced gradient
    assert got[0]["_loss"].item() == pytest.approx(total.item(), rel=1e-5)


def _bucket_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        red = GradAllReduce(torch.device("cpu"))              # "auto": no CUDA here -> plain bucket, gloo reduction
        buf = red.alloc(1001)
        assert buf.shape == (1001,) and not red.uses_nvls(buf) and float(buf.abs().sum()) == 0.0
        buf += torch.arange(1001, dtype=torch.float32) * (rank + 1)
        red.launch([buf])
        red.wait()
        torch.save(buf, os.path.join(out_dir, f"bucket{rank}.pt"))
        with pytest.raises(RuntimeError):
            GradAllReduce(torch.device("cpu"), backend="nvls").alloc(16)   # the in-switch kernel needs CUDA buckets
    finally:
        dist.destroy_process_group()


def test_bucket_alloc_falls_back_to_the_process_group_without_nvswitch(tmp_path):
    world = 2
    mp.spawn(_bucket_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    want = torch.arange(1001, dtype=torch.float32) * 3
    for r in range(world):
        assert torch.equal(torch.load(os.path.join(str(tmp_path), f"bucket{r}.pt")), want)
    with pytest.raises(ValueError):
        GradAllReduce(torch.device("cpu"), backend="rdma")


def test_numa_binding_helper_is_a_safe_no_op_without_topology():
    from spfsplatv2_b200.dp import _parse_cpulist, bind_to_gpu_numa_node
    assert _parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert _parse_cpulist("") == []
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa_node(0) is None or isinstance(bind_to_gpu_numa_node(0), int)   # never raises
    if not torch.cuda.is_available():
        assert os.sched_getaffinity(0) == before


def _ddp_worker(rank, world, port, out_dir, use_hook):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from torch.nn.parallel import DistributedDataParallel as DDP
        from spfsplatv2_b200.dp import nvls_comm_hook
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(24, 64), torch.nn.GELU(), torch.nn.Linear(64, 64), torch.nn.GELU(),
                                    torch.nn.Linear(64, 3))
        ddp = DDP(model, bucket_cap_mb=0.01, find_unused_parameters=True)        # several small buckets, as the reference's strategy
        if use_hook:
            ddp.register_comm_hook(None, nvls_comm_hook(GradAllReduce(torch.device("cpu"))))
        g = torch.Generator().manual_seed(100 + rank)                            # different data per rank
        for step in range(3):                                                    # twins are reused from the second step on
            x = torch.randn(16, 24, generator=g)
            ddp.zero_grad()
            ddp(x).square().mean().backward()
        torch.save([p.grad.clone() for p in model.parameters()], os.path.join(out_dir, f"ddp{int(use_hook)}_{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_ddp_comm_hook_reproduces_ddps_own_allreduce(tmp_path):
    """spfsplatv2_b200.dp.nvls_comm_hook registered on torch DDP (src/main.py:141-145: the reference trains under DDP):
    gradients equal those of DDP's built-in all-reduce, on every rank, over several steps and several buckets.  On CPU
    the reducer's buckets are plain tensors reduced by gloo -- the hook's plumbing (twin allocation per bucket, copy in,
    reduce, average, copy out, future) is what is exercised; the in-switch kernel is checked by bench.py --gpus N."""
    world = 2
    for use_hook in (False, True):
        mp.spawn(_ddp_worker, args=(world, _free_port(), str(tmp_path), use_hook), nprocs=world, join=True)
    for r in range(world):
        a = torch.load(os.path.join(str(tmp_path), f"ddp0_{r}.pt"))
        b = torch.load(os.path.join(str(tmp_path), f"ddp1_{r}.pt"))
        assert len(a) == len(b) == 6
        for x, y in zip(a, b):
            assert torch.allclose(x, y, rtol=1e-6, atol=1e-8)
    a0 = torch.load(os.path.join(str(tmp_path), "ddp1_0.pt"))
    a1 = torch.load(os.path.join(str(tmp_path), "ddp1_1.pt"))
    for x, y in zip(a0, a1):
        assert torch.equal(x, y)

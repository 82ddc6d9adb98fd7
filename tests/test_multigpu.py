"""Multi-GPU host logic. CPU: world_size-2 gloo processes exercise the sharding and the band stitch with CPU tensors.
GPU (needs >= 2 devices, skipped otherwise): one huge grid decoded as tile-row bands on 2 GPUs and stitched over NCCL
equals the single-GPU decode (BASELINE config C5)."""
import hashlib
import os
import socket
import sys

import numpy as np
import pytest

from conftest import PKG_DIR, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_files_is_a_balanced_partition():
    from heif_b200.multigpu import shard_files
    sizes = [5, 1, 9, 3, 3, 7, 2, 8, 8, 1, 4]
    for world in (1, 2, 3, 4, 8):
        parts = [shard_files(sizes, world, r) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(len(sizes)))
        loads = [sum(sizes[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(sizes)


def test_tile_row_bands_cover_the_grid():
    from heif_b200.multigpu import tile_row_band
    for rows in (1, 6, 17, 64):
        for world in (1, 2, 4, 8):
            bands = [tile_row_band(rows, world, r) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == rows
            assert all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
            lens = [e - b for b, e in bands]
            assert max(lens) - min(lens) <= 1


def _gloo_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, PKG_DIR)
    import torch
    import torch.distributed as dist
    from heif_b200.multigpu import shard_files, stitch_bands, tile_row_band
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # a fake 7-tile-row image, 5 output rows per tile row except the last (3): rank r owns band r
    rows_per_tile, tile_rows, row_bytes, full_rows = 5, 7, 24, 33
    b, e = tile_row_band(tile_rows, world, rank)
    y0, y1 = b * rows_per_tile, min(full_rows, e * rows_per_tile)
    whole = (torch.arange(full_rows * row_bytes, dtype=torch.int64) % 251).to(torch.uint8).reshape(full_rows, row_bytes)
    full = stitch_bands(whole[y0:y1].clone(), y0, full_rows, dist, rank, world)
    ok = (full is None) if rank else bool(torch.equal(full, whole))
    # image sharding: every rank computes the same partition
    mine = shard_files([3, 9, 4, 4, 1, 7], world, rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ok = ok and sorted(i for p in gathered for i in p) == list(range(6))
    open(os.path.join(out_dir, "rank%d" % rank), "w").write("ok" if ok else "bad")
    dist.destroy_process_group()


def test_gloo_world2_stitch_and_sharding(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert [open(os.path.join(tmp_path, "rank%d" % r)).read() for r in range(2)] == ["ok", "ok"]


def _shared_worker(rank, world, port, out_dir, path):
    """one process per GPU; the bands are written into rank 0's buffer by K5 itself (CUDA IPC + NVLink peer stores)"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, PKG_DIR)
    import torch
    import torch.distributed as dist
    import heif_b200 as hb
    from heif_b200.multigpu import decode_grid_shared
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)      # only the IPC handle and barriers travel here
    data = open(path, "rb").read()
    eng = hb.Engine(rank)
    full = decode_grid_shared(eng, data, dist, rank, world)
    if rank == 0:
        open(os.path.join(out_dir, "md5_shared"), "w").write(hashlib.md5(full.tobytes()).hexdigest())
    eng.close()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_gpu_bands_into_a_shared_image_same_process(tmp_path):
    """C5 behind the C ABI on whatever GPUs this box has: two engines (on two devices when there are two, else both on
    device 0) decode the two halves of a grid; both K5 kernels write into ONE hc_shared_image of the first engine
    (peer stores through hc_shared_image_attach when the devices differ). Equals the single-job decode."""
    import ctypes as C
    import torch
    import heif_b200 as hb
    from heif_b200.multigpu import BandJob, SharedImage
    sys.path.insert(0, ROOT)
    from tools import heif_writer
    data = heif_writer.synth_grid_heic(1280, 1100, tile=256, seed=9, qp=30, wpp=1, sao=1, log2_ctb=5)   # 5 tile rows, last one cropped
    ndev = torch.cuda.device_count()
    engines = [hb.Engine(0), hb.Engine(1 if ndev > 1 else 0)]
    want = hb.decode_heic(engines[0], data, hb.OUT_RGB)
    jobs = [BandJob(engines[r], data, r, 2) for r in range(2)]
    owner = SharedImage(engines[0], jobs[0].width, jobs[0].full_height, 3)
    views = [owner, SharedImage.attach(engines[1], owner)]
    L = engines[0]._L
    for r in range(2):
        hb.api.check(L, L.hc_heic_job_set_rgb_target(jobs[r]._h, 0, views[r]._h, jobs[r].first_row), "set_rgb_target")
        hb.api.check(L, L.hc_heic_job_upload(jobs[r]._h), "upload")
        hb.api.check(L, L.hc_heic_job_run(jobs[r]._h), "run")
    for r in range(2):
        hb.api.check(L, L.hc_heic_job_sync(jobs[r]._h), "sync")
    got = owner.read()
    assert got.shape == want.shape and np.array_equal(got, want)
    # an image with an external target is read there, not through the job
    with pytest.raises(hb.HeifCudaError):
        out = np.empty((jobs[0].desc.height, jobs[0].width * 3), np.uint8)
        hb.api.check(L, L.hc_heic_job_read_rgb(jobs[0]._h, 0, out.ctypes.data, out.strides[0]), "read_rgb")
    for j in jobs:
        j.close()
    views[1].close()
    owner.close()
    for e in engines:
        e.close()


@pytest.mark.gpu
def test_gpu_grid_bands_peer_written_on_two_gpus_match_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import heif_b200 as hb
    sys.path.insert(0, ROOT)
    from tools import heif_writer
    data = heif_writer.synth_grid_heic(1280, 1100, tile=256, seed=9, qp=30, wpp=1, sao=1, log2_ctb=5)
    path = os.path.join(tmp_path, "grid.heic")
    open(path, "wb").write(data)
    eng = hb.Engine(0)
    want = hashlib.md5(hb.decode_heic(eng, data, hb.OUT_RGB).tobytes()).hexdigest()
    eng.close()
    mp.spawn(_shared_worker, args=(2, _free_port(), str(tmp_path), path), nprocs=2, join=True)
    assert open(os.path.join(tmp_path, "md5_shared")).read() == want


def _nccl_worker(rank, world, port, out_dir, path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, PKG_DIR)
    import torch
    import torch.distributed as dist
    import heif_b200 as hb
    from heif_b200.multigpu import decode_grid_sharded
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    data = open(path, "rb").read()
    eng = hb.Engine(rank)
    full = decode_grid_sharded(eng, data, dist, rank, world, torch.device("cuda", rank))
    if rank == 0:
        open(os.path.join(out_dir, "md5"), "w").write(hashlib.md5(full.cpu().numpy().tobytes()).hexdigest())
    torch.cuda.synchronize()
    eng.close()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_gpu_grid_bands_on_two_gpus_match_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import heif_b200 as hb
    sys.path.insert(0, ROOT)
    from tools import heif_writer
    data = heif_writer.synth_grid_heic(1280, 1100, tile=256, seed=9, qp=30, wpp=1, sao=1, log2_ctb=5)   # 5 tile rows, last one cropped
    path = os.path.join(tmp_path, "grid.heic")
    open(path, "wb").write(data)
    eng = hb.Engine(0)
    want = hashlib.md5(hb.decode_heic(eng, data, hb.OUT_RGB).tobytes()).hexdigest()
    eng.close()
    mp.spawn(_nccl_worker, args=(2, _free_port(), str(tmp_path), path), nprocs=2, join=True)
    assert open(os.path.join(tmp_path, "md5")).read() == want

"""Multi-GPU host logic: one process per GPU (torch.distributed for the launch plumbing), no data-path collective.

The reference parallelises over HEIF grid tiles with std::async inside one process
(libheif/context.cc:2361-2401) and over pictures not at all. Here
  * a list of files is sharded by image across ranks (BASELINE config C4) — zero inter-GPU traffic;
  * ONE huge grid image is cut into bands of tile rows, one band per rank (config C5); every rank runs K0..K5 on its
    tiles and its K5 stores the band's RGB rows STRAIGHT into the owner GPU's output buffer (hc_shared_image behind the C
    ABI: CUDA IPC mapping of the owner's allocation, NVLink peer stores) — decode_grid_shared. torch.distributed only
    carries the 64-byte IPC handle and the barrier. Grid tiles are independent HEVC pictures (context.cc:2407-2415):
    nothing crosses GPUs before that write. stitch_bands (point-to-point send / recv of whole bands) remains for backends
    without peer memory (gloo in the CPU tests).
"""
import ctypes as C

from ._lib import HeifCudaError, ImageDesc, check


def shard_files(sizes, world, rank):
    """Indices of the files rank `rank` decodes: greedy size-balanced assignment (largest first), deterministic on
    every rank. `sizes` are byte sizes (a proxy for CABAC work, which is what limits the host side)."""
    load = [0] * world
    mine = []
    for i in sorted(range(len(sizes)), key=lambda k: (-sizes[k], k)):
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += sizes[i]
        if r == rank:
            mine.append(i)
    return sorted(mine)


def tile_row_band(rows, world, rank):
    """Contiguous band [begin, end) of the grid's tile rows for this rank (empty when rows < world for high ranks)."""
    base, extra = divmod(rows, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def stitch_bands(band, first_row, full_rows, dist, rank, world, owner=0, group=None):
    """Gathers row bands into the owner's full image with point-to-point copies.
    band: 2-D uint8 tensor [rows_of_this_rank, row_bytes] (may have 0 rows); every rank passes the first output row
    of its band. Returns the full [full_rows, row_bytes] tensor on the owner, None elsewhere."""
    import torch
    meta = torch.tensor([first_row, band.shape[0]], dtype=torch.int64, device=band.device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)          # 16 bytes per rank: where does every band go
    if rank != owner:
        if band.shape[0]:
            dist.send(band.contiguous(), dst=owner, group=group)
        return None
    full = torch.empty((full_rows, band.shape[1]), dtype=torch.uint8, device=band.device)
    reqs = []
    for r in range(world):
        y0, n = int(metas[r][0]), int(metas[r][1])
        if n == 0:
            continue
        if r == owner:
            full[y0:y0 + n].copy_(band)
        else:
            reqs.append(dist.irecv(full[y0:y0 + n], src=r, group=group))   # whole rows: a contiguous slice, no staging
    for q in reqs:
        q.wait()
    return full


class BandJob:
    """hc_heic_job_create_band: this rank's band of tile rows of one grid image."""

    def __init__(self, engine, data, rank, world, want_alpha=False, threads=0):
        from .api import HeifFile
        self._L = engine._L
        hf = HeifFile(data)
        info = hf.image_info(hf.primary_id)
        hf.close()
        if not info.is_grid:
            raise HeifCudaError("banded decode needs a grid image")
        self.begin, self.end = tile_row_band(info.rows, world, rank)
        self._h = None
        self.first_row, self.full_height, self.desc = 0, info.height, None
        self.width = info.width
        if self.end > self.begin:
            self._buf = C.create_string_buffer(data, len(data))
            y0, fh = C.c_int(0), C.c_int(0)
            self._h = self._L.hc_heic_job_create_band(engine._h, C.cast(self._buf, C.c_char_p), len(data), int(want_alpha), threads,
                                                      self.begin, self.end, C.byref(y0), C.byref(fh))
            if not self._h:
                raise HeifCudaError("band job: " + (self._L.hc_last_error() or b"").decode())
            self.first_row, self.full_height = y0.value, fh.value
            self.desc = ImageDesc()
            check(self._L, self._L.hc_heic_job_image_desc(self._h, 0, C.byref(self.desc)), "image_desc")

    def run_into(self, tensor):
        """upload + K1..K5, then device-to-device copy of the band into `tensor` ([band_rows, row_bytes] uint8, CUDA)."""
        if not self._h:
            return
        check(self._L, self._L.hc_heic_job_upload(self._h), "upload")
        check(self._L, self._L.hc_heic_job_run(self._h), "run")
        check(self._L, self._L.hc_heic_job_copy_rgb_device(self._h, 0, C.c_void_p(tensor.data_ptr()), tensor.stride(0)), "copy_rgb_device")

    def stage_ms(self):
        ms = (C.c_float * 8)()
        check(self._L, self._L.hc_heic_job_stage_ms(self._h, ms), "stage_ms")
        return dict(zip(("h2d", "k1_transform", "k2_intra", "k3_deblock", "k4_sao", "k5_csc", "d2h", "k0_parse"), list(ms)[:8]))

    def close(self):
        if self._h:
            self._L.hc_heic_job_destroy(self._h)
            self._h = None


def decode_grid_sharded(engine, data, dist, rank, world, device, want_alpha=False, threads=0, owner=0):
    """One huge grid HEIC decoded by all ranks; the owner returns the whole interleaved image as a CUDA tensor."""
    import torch
    job = BandJob(engine, data, rank, world, want_alpha, threads)
    try:
        bpp = job.desc.bytes_per_pixel if job.desc is not None else (4 if want_alpha else 3)
        rows = job.desc.height if job.desc is not None else 0
        band = torch.empty((rows, job.width * bpp), dtype=torch.uint8, device=device)
        job.run_into(band)
        return stitch_bands(band, job.first_row, job.full_height, dist, rank, world, owner)
    finally:
        job.close()


class SharedImage:
    """hc_shared_image: one RGB buffer on the owner GPU that the K5 kernels of other GPUs / processes store into."""

    def __init__(self, engine, width, height, bytes_per_pixel, handle=None):
        self._L = engine._L
        self.width, self.height, self.bpp = width, height, bytes_per_pixel
        if handle is None:
            self._h = self._L.hc_shared_image_create(engine._h, width, height, bytes_per_pixel)
        else:
            buf = (C.c_uint8 * 64)(*handle)
            self._h = self._L.hc_shared_image_open(engine._h, buf, width, height, bytes_per_pixel)
        if not self._h:
            raise HeifCudaError("shared image: " + (self._L.hc_last_error() or b"").decode())

    @classmethod
    def attach(cls, engine, owner):
        """view of `owner` (same process) for an engine on another device: enables peer access"""
        s = cls.__new__(cls)
        s._L, s.width, s.height, s.bpp = engine._L, owner.width, owner.height, owner.bpp
        s._h = s._L.hc_shared_image_attach(engine._h, owner._h)
        if not s._h:
            raise HeifCudaError("shared image attach: " + (s._L.hc_last_error() or b"").decode())
        return s

    def export(self):
        buf = (C.c_uint8 * 64)()
        check(self._L, self._L.hc_shared_image_export(self._h, buf), "hc_shared_image_export")
        return bytes(buf)

    def read(self, first_row=0, rows=None):
        import numpy as np
        rows = self.height - first_row if rows is None else rows
        out = np.empty((rows, self.width * self.bpp), np.uint8)
        check(self._L, self._L.hc_shared_image_read(self._h, first_row, rows, out.ctypes.data, out.strides[0]), "hc_shared_image_read")
        return out

    def close(self):
        if self._h:
            self._L.hc_shared_image_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


def decode_grid_shared(engine, data, dist, rank, world, want_alpha=False, threads=0, owner=0, out_format=None, timings=None):
    """One huge grid HEIC decoded by all ranks (one process per GPU); every rank's K5 writes its band into the owner's
    buffer over NVLink. Returns the whole interleaved image as a numpy array on the owner, None elsewhere.
    `dist`: an initialised torch.distributed (any backend) — used for one broadcast of the IPC handle and two barriers."""
    import time
    import torch
    sel = (0x100 | int(out_format)) if out_format is not None else int(want_alpha)
    job = BandJob(engine, data, rank, world, sel, threads)
    shared = None
    try:
        # every rank knows the full geometry from the container; pixel size from its own band (ranks without tiles ask rank 0's)
        meta = [job.width, job.full_height, job.desc.bytes_per_pixel if job.desc is not None else 0]
        metas = [None] * world
        dist.all_gather_object(metas, meta)
        bpp = max(m[2] for m in metas)
        width, height = metas[owner][0], metas[owner][1]
        handle = [None]
        if rank == owner:
            shared = SharedImage(engine, width, height, bpp)
            handle[0] = shared.export()
        dist.broadcast_object_list(handle, src=owner)
        if rank != owner:
            shared = SharedImage(engine, width, height, bpp, handle=handle[0])
        dist.barrier()
        t0 = time.perf_counter()
        if job._h:
            check(job._L, job._L.hc_heic_job_set_rgb_target(job._h, 0, shared._h, job.first_row), "hc_heic_job_set_rgb_target")
            check(job._L, job._L.hc_heic_job_upload(job._h), "upload")
            check(job._L, job._L.hc_heic_job_run(job._h), "run")
            check(job._L, job._L.hc_heic_job_sync(job._h), "sync")
        t1 = time.perf_counter()
        dist.barrier()                       # every band is in the owner's memory
        t2 = time.perf_counter()
        if timings is not None:
            timings.update(run_s=t1 - t0, total_s=t2 - t0, stage_ms=job.stage_ms() if job._h else {})
        out = shared.read() if rank == owner else None
        dist.barrier()                       # the owner has read: peers may unmap
        return out
    finally:
        if shared is not None:
            shared.close()
        job.close()

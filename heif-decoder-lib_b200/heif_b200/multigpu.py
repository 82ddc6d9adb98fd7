"""Multi-GPU host logic: one process per GPU (torch.distributed), no data-path collective.

The reference parallelises over HEIF grid tiles with std::async inside one process
(libheif/context.cc:2361-2401) and over pictures not at all. Here
  * a list of files is sharded by image across ranks (BASELINE config C4) — zero inter-GPU traffic;
  * ONE huge grid image is cut into bands of tile rows, one band per rank (config C5); every rank runs K1..K5 on its
    tiles and the bands are stitched into the owner rank's RGB buffer with point-to-point copies (NCCL send/recv over
    NVLink on GPUs, gloo in the CPU tests). Grid tiles are independent HEVC pictures (context.cc:2407-2415): nothing
    crosses GPUs before the stitch.
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

"""Minimal HEIF (ISO/IEC 23008-12) writer for test and benchmark content: wraps HEVC intra streams
from tools/hevcenc.py (or any length-prefixed NAL stream) into .heic files — single images, grids of
tiles (like iPhone photos) and images with an alpha auxiliary image. Test infrastructure only.

The files it writes are read by the unmodified reference (oracle/_ref/libheifref.so) in the tests,
which is what validates this writer."""
import struct

from . import hevcenc


def box(fourcc, payload):
    return struct.pack(">I4s", 8 + len(payload), fourcc) + payload


def fullbox(fourcc, version, flags, payload):
    return box(fourcc, struct.pack(">I", (version << 24) | flags) + payload)


def hvcc_from_stream(stream, chroma_format, bit_depth):
    """hvcC property + the slice NALs (length-prefixed) of one coded picture."""
    nals = hevcenc.split_nals(stream)
    ps = [n for n in nals if ((n[0] >> 1) & 0x3F) in (32, 33, 34)]
    slices = [n for n in nals if ((n[0] >> 1) & 0x3F) < 32]
    profile = 1 if (chroma_format == 1 and bit_depth == 8) else (2 if chroma_format == 1 and bit_depth <= 10 else 4)
    cfg = bytes([1, profile]) + struct.pack(">I", 1 << (31 - profile)) + bytes(6) + bytes([186])
    cfg += struct.pack(">H", 0xF000) + bytes([0xFC, 0xFC | chroma_format, 0xF8 | (bit_depth - 8), 0xF8 | (bit_depth - 8)])
    cfg += struct.pack(">H", 0) + bytes([(1 << 3) | (1 << 2) | 3])
    arrays = b""
    types = sorted(set((n[0] >> 1) & 0x3F for n in ps))
    for t in types:
        group = [n for n in ps if ((n[0] >> 1) & 0x3F) == t]
        arrays += bytes([0x80 | t]) + struct.pack(">H", len(group))
        for n in group:
            arrays += struct.pack(">H", len(n)) + n
    cfg += bytes([len(types)]) + arrays
    data = b"".join(struct.pack(">I", len(n)) + n for n in slices)
    return box(b"hvcC", cfg), data


class HeifBuilder:
    def __init__(self):
        self.items = []      # dict(id, type, data, hidden, props[list of prop indices])
        self.props = []      # property boxes (bytes)
        self.refs = []       # (type, from, [to])
        self.primary = None

    def add_prop(self, b):
        if b in self.props:
            return self.props.index(b) + 1
        self.props.append(b)
        return len(self.props)

    def add_item(self, item_type, data, props, hidden=False):
        iid = len(self.items) + 1
        self.items.append(dict(id=iid, type=item_type, data=data, hidden=hidden, props=props))
        return iid

    def add_hevc_image(self, stream, width, height, chroma_format, bit_depth, hidden=False, nclx=None, extra_props=()):
        hvcc, data = hvcc_from_stream(stream, chroma_format, bit_depth)
        props = [self.add_prop(hvcc), self.add_prop(fullbox(b"ispe", 0, 0, struct.pack(">II", width, height)))]
        if nclx is not None:
            prim, trc, mat, full = nclx
            props.append(self.add_prop(box(b"colr", b"nclx" + struct.pack(">HHHB", prim, trc, mat, 0x80 if full else 0))))
        for p in extra_props:
            props.append(self.add_prop(p))
        return self.add_item(b"hvc1", data, props, hidden)

    def add_grid(self, tile_ids, rows, cols, out_w, out_h, nclx=None, extra_props=()):
        big = out_w > 65535 or out_h > 65535
        payload = bytes([0, 1 if big else 0, rows - 1, cols - 1]) + (struct.pack(">II", out_w, out_h) if big else struct.pack(">HH", out_w, out_h))
        props = [self.add_prop(fullbox(b"ispe", 0, 0, struct.pack(">II", out_w, out_h)))]
        if nclx is not None:
            prim, trc, mat, full = nclx
            props.append(self.add_prop(box(b"colr", b"nclx" + struct.pack(">HHHB", prim, trc, mat, 0x80 if full else 0))))
        for p in extra_props:
            props.append(self.add_prop(p))
        gid = self.add_item(b"grid", payload, props)
        self.refs.append((b"dimg", gid, list(tile_ids)))
        return gid

    def add_alpha_grid(self, master_id, tile_streams, rows, cols, tile_w, tile_h, out_w, out_h, bit_depth, chroma_format=0,
                       premultiplied=False):
        """alpha auxiliary image that is itself a 'grid' of monochrome HEVC tiles"""
        auxc = fullbox(b"auxC", 0, 0, b"urn:mpeg:hevc:2015:auxid:1\x00")
        tiles = [self.add_hevc_image(t, tile_w, tile_h, chroma_format, bit_depth, hidden=True) for t in tile_streams]
        aid = self.add_grid(tiles, rows, cols, out_w, out_h, extra_props=(auxc,))
        self.items[aid - 1]["hidden"] = True
        self.refs.append((b"auxl", aid, [master_id]))
        if premultiplied:
            self.refs.append((b"prem", master_id, [aid]))
        return aid

    def add_overlay(self, child_ids, canvas_w, canvas_h, offsets, background=(0, 0, 0, 0xffff), extra_props=()):
        """'iovl' derived image (ISO/IEC 23008-12 6.6.2.2; libheif context.cc:318-369): 16-bit RGBA background colour,
        canvas size and one signed (x, y) offset per referenced image"""
        big = canvas_w > 65535 or canvas_h > 65535 or any(abs(v) > 32767 for o in offsets for v in o)
        f = ">II" if big else ">HH"
        fs = ">ii" if big else ">hh"
        payload = bytes([0, 1 if big else 0]) + struct.pack(">HHHH", *background) + struct.pack(f, canvas_w, canvas_h)
        payload += b"".join(struct.pack(fs, x, y) for x, y in offsets)
        props = [self.add_prop(fullbox(b"ispe", 0, 0, struct.pack(">II", canvas_w, canvas_h)))]
        for p in extra_props:
            props.append(self.add_prop(p))
        oid = self.add_item(b"iovl", payload, props)
        self.refs.append((b"dimg", oid, list(child_ids)))
        return oid

    def add_alpha(self, master_id, stream, width, height, chroma_format, bit_depth, extra_props=(), premultiplied=False):
        auxc = fullbox(b"auxC", 0, 0, b"urn:mpeg:hevc:2015:auxid:1\x00")
        aid = self.add_hevc_image(stream, width, height, chroma_format, bit_depth, hidden=True, extra_props=(auxc,) + tuple(extra_props))
        self.refs.append((b"auxl", aid, [master_id]))
        if premultiplied:       # the colour samples are stored premultiplied by this alpha image (libheif context.cc:1150-1161)
            self.refs.append((b"prem", master_id, [aid]))
        return aid

    def serialize(self):
        primary = self.primary or self.items[0]["id"]
        ftyp = box(b"ftyp", b"heic" + struct.pack(">I", 0) + b"mif1heic")
        hdlr = fullbox(b"hdlr", 0, 0, struct.pack(">I4s", 0, b"pict") + bytes(12) + b"\x00")
        pitm = fullbox(b"pitm", 0, 0, struct.pack(">H", primary))
        infes = b""
        for it in self.items:
            infes += fullbox(b"infe", 2, 1 if it["hidden"] else 0, struct.pack(">HH4s", it["id"], 0, it["type"]) + b"\x00")
        iinf = fullbox(b"iinf", 0, 0, struct.pack(">H", len(self.items)) + infes)
        iref = b""
        if self.refs:
            body = b""
            for t, frm, to in self.refs:
                body += box(t, struct.pack(">HH", frm, len(to)) + b"".join(struct.pack(">H", x) for x in to))
            iref = fullbox(b"iref", 0, 0, body)
        ipco = box(b"ipco", b"".join(self.props))
        ipma_body = struct.pack(">I", len(self.items))
        for it in self.items:
            ipma_body += struct.pack(">HB", it["id"], len(it["props"])) + bytes((0x80 | p) for p in it["props"])
        iprp = box(b"iprp", ipco + fullbox(b"ipma", 0, 0, ipma_body))

        def build_meta(offsets):
            iloc_body = struct.pack(">BBH", 0x44, 0x00, len(self.items))
            for it, off in zip(self.items, offsets):
                iloc_body += struct.pack(">HHHH", it["id"], 0, 0, 1) + struct.pack(">II", off, len(it["data"]))
            iloc = fullbox(b"iloc", 1, 0, iloc_body)
            return fullbox(b"meta", 0, 0, hdlr + pitm + iloc + iinf + iref + iprp)

        meta0 = build_meta([0] * len(self.items))
        base = len(ftyp) + len(meta0) + 8
        offsets, pos = [], base
        for it in self.items:
            offsets.append(pos)
            pos += len(it["data"])
        meta = build_meta(offsets)
        assert len(meta) == len(meta0)
        mdat = box(b"mdat", b"".join(it["data"] for it in self.items))
        return ftyp + meta + mdat


def irot(quarter_turns_ccw):
    """ImageRotation property (anti-clockwise quarter turns)"""
    return box(b"irot", bytes([quarter_turns_ccw & 3]))


def imir(axis):
    """ImageMirror property: axis bit 0 (libheif box.cc:3626: 1 -> "horizontal" direction, every row reversed)"""
    return box(b"imir", bytes([axis & 1]))


def clap(w_num, w_den, h_num, h_den, hoff_num, hoff_den, voff_num, voff_den):
    """CleanAperture property"""
    return box(b"clap", struct.pack(">IIIIiIiI", w_num, w_den, h_num, h_den, hoff_num, hoff_den, voff_num, voff_den))


def single_image(stream, width, height, chroma_format=1, bit_depth=8, nclx=None, alpha_stream=None, alpha_chroma_format=0,
                 transforms=(), alpha_size=None, alpha_transforms=None, premultiplied=False):
    """transforms: property boxes (irot / imir / clap) attached, in this order, to the image and to its alpha image
    (alpha_transforms: the alpha image's own list instead); alpha_size: (width, height) of the alpha image when it
    differs from the colour image's (the reference rescales it by nearest neighbour, context.cc:2064-2071)"""
    b = HeifBuilder()
    iid = b.add_hevc_image(stream, width, height, chroma_format, bit_depth, nclx=nclx, extra_props=tuple(transforms))
    if alpha_stream is not None:
        aw, ah = alpha_size if alpha_size else (width, height)
        b.add_alpha(iid, alpha_stream, aw, ah, alpha_chroma_format, bit_depth,
                    extra_props=tuple(transforms if alpha_transforms is None else alpha_transforms), premultiplied=premultiplied)
    b.primary = iid
    return b.serialize()


def grid_image(tile_streams, rows, cols, tile_w, tile_h, out_w, out_h, chroma_format=1, bit_depth=8, nclx=None,
               tile_nclx=None, transforms=()):
    b = HeifBuilder()
    tiles = [b.add_hevc_image(s, tile_w, tile_h, chroma_format, bit_depth, hidden=True, nclx=tile_nclx) for s in tile_streams]
    gid = b.add_grid(tiles, rows, cols, out_w, out_h, nclx=nclx, extra_props=tuple(transforms))
    b.primary = gid
    return b.serialize()


def synth_grid_heic(out_w, out_h, tile=512, chroma_format=1, bit_depth=8, seed=0, transforms=(), **enc_opts):
    """A whole synthetic photo cut into tile x tile HEVC pictures, iPhone style (BASELINE config C2)."""
    cols, rows = (out_w + tile - 1) // tile, (out_h + tile - 1) // tile
    full = hevcenc.synth_image(cols * tile, rows * tile, chroma_format, bit_depth, seed)
    sw = 2 if chroma_format in (1, 2) else 1
    sh = 2 if chroma_format == 1 else 1
    streams = []
    for r in range(rows):
        for c in range(cols):
            planes = [full[0][r * tile:(r + 1) * tile, c * tile:(c + 1) * tile]]
            if chroma_format:
                for k in (1, 2):
                    planes.append(full[k][r * tile // sh:(r + 1) * tile // sh, c * tile // sw:(c + 1) * tile // sw])
            opts = dict(enc_opts)
            opts.setdefault("seed", seed * 1000 + r * cols + c + 1)
            streams.append(hevcenc.encode(planes, chroma_format=chroma_format, bit_depth=bit_depth, **opts))
    return grid_image(streams, rows, cols, tile, tile, out_w, out_h, chroma_format, bit_depth, transforms=transforms)

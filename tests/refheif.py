"""ctypes binding to the UNMODIFIED reference libheif built by oracle/Makefile.ref (oracle/_ref/libheifref.so).

Test infrastructure only (the checker): used by tests/ and bench.py's reference arm.
API used: libheif/api/libheif/heif.h  (heif_context_alloc :905, heif_context_read_from_memory_without_copy,
heif_context_get_primary_image_handle, heif_decode_image :1634, heif_image_get_plane_readonly :1731).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(os.path.dirname(_HERE), "oracle", "_ref")

# enum values: heif.h heif_colorspace / heif_chroma / heif_channel
COLORSPACE_UNDEFINED, COLORSPACE_YCBCR, COLORSPACE_RGB, COLORSPACE_MONO = 99, 0, 1, 2
CHROMA_UNDEFINED, CHROMA_MONO, CHROMA_420, CHROMA_422, CHROMA_444 = 99, 0, 1, 2, 3
CHROMA_RGB, CHROMA_RGBA, CHROMA_RRGGBB_BE, CHROMA_RRGGBBAA_BE, CHROMA_RRGGBB_LE, CHROMA_RRGGBBAA_LE = 10, 11, 12, 13, 14, 15
CH_Y, CH_CB, CH_CR, CH_R, CH_G, CH_B, CH_ALPHA, CH_INTERLEAVED = 0, 1, 2, 3, 4, 5, 6, 10


class HeifError(C.Structure):
    _fields_ = [("code", C.c_int), ("subcode", C.c_int), ("message", C.c_char_p)]


_lib = None
# set by tests/test_plugin.py once libheif-cuda.so is registered in this process: from then on a
# plain decode() must name the reference's own decoder explicitly to stay the reference
FOREIGN_PLUGIN_LOADED = False


def available():
    return os.path.exists(os.path.join(REF_DIR, "libheifref.so"))


def lib():
    global _lib
    if _lib is None:
        # RTLD_LOCAL on purpose (libde265ref.so comes in through libheifref.so's $ORIGIN rpath): the reference's
        # generic C++ symbol names must not interpose other libraries of the test process (torch crashed on import)
        L = C.CDLL(os.path.join(REF_DIR, "libheifref.so"))
        L.heif_context_alloc.restype = C.c_void_p
        L.heif_context_free.argtypes = [C.c_void_p]
        L.heif_context_read_from_memory_without_copy.restype = HeifError
        L.heif_context_read_from_memory_without_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.heif_context_get_primary_image_handle.restype = HeifError
        L.heif_context_get_primary_image_handle.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.heif_context_get_image_handle.restype = HeifError
        L.heif_context_get_image_handle.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]
        L.heif_context_get_number_of_top_level_images.argtypes = [C.c_void_p]
        L.heif_context_get_list_of_top_level_image_IDs.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_int]
        L.heif_context_set_threads.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.heif_context_set_threads.restype = None
        L.heif_image_handle_release.argtypes = [C.c_void_p]
        L.heif_image_handle_get_width.argtypes = [C.c_void_p]
        L.heif_image_handle_get_height.argtypes = [C.c_void_p]
        L.heif_decode_image.restype = HeifError
        L.heif_decode_image.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p]
        L.heif_image_release.argtypes = [C.c_void_p]
        L.heif_image_get_plane_readonly.restype = C.POINTER(C.c_uint8)
        L.heif_image_get_plane_readonly.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.heif_image_get_width.argtypes = [C.c_void_p, C.c_int]
        L.heif_image_get_height.argtypes = [C.c_void_p, C.c_int]
        L.heif_image_get_bits_per_pixel_range.argtypes = [C.c_void_p, C.c_int]
        L.heif_image_get_chroma_format.argtypes = [C.c_void_p]
        L.heif_image_get_colorspace.argtypes = [C.c_void_p]
        L.heif_image_has_channel.argtypes = [C.c_void_p, C.c_int]
        L.heif_load_plugin.restype = HeifError
        L.heif_load_plugin.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.heif_decoding_options_alloc.restype = C.c_void_p
        L.heif_decoding_options_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _check(err, what):
    if err.code != 0:
        raise RuntimeError("%s: heif_error %d/%d %s" % (what, err.code, err.subcode, err.message))


def plane_bytes(img, channel, bytes_per_px):
    """Tightly packed rows of one channel as bytes."""
    L = lib()
    stride = C.c_int(0)
    p = L.heif_image_get_plane_readonly(img, channel, C.byref(stride))
    if not p:
        return None, 0, 0
    w = L.heif_image_get_width(img, channel)
    h = L.heif_image_get_height(img, channel)
    row = w * bytes_per_px
    buf = C.string_at(p, stride.value * (h - 1) + row)
    if stride.value == row:
        return buf, w, h
    return b"".join(buf[y * stride.value: y * stride.value + row] for y in range(h)), w, h


def decode(data, colorspace, chroma, item_id=None, threads=None, decoder_id=None, bilinear=False, premultiply=False):
    """heif_decode_image on an in-memory HEIC. Returns dict of channel -> (bytes, w, h) plus bpp.
    decoder_id: None = the reference's libde265 plugin; "" = let libheif choose by priority."""
    L = lib()
    if decoder_id is None and FOREIGN_PLUGIN_LOADED:
        decoder_id = "libde265"
    if decoder_id == "":
        decoder_id = None
    ctx = L.heif_context_alloc()
    buf = C.create_string_buffer(data, len(data))
    try:
        _check(L.heif_context_read_from_memory_without_copy(ctx, buf, len(data), None), "read")
        h = C.c_void_p()
        if item_id is None:
            _check(L.heif_context_get_primary_image_handle(ctx, C.byref(h)), "primary handle")
        else:
            _check(L.heif_context_get_image_handle(ctx, item_id, C.byref(h)), "handle")
        if threads is not None:
            L.heif_context_set_threads(ctx, h, threads)
        img = C.c_void_p()
        opts = None
        if decoder_id is not None or bilinear:
            opts = L.heif_decoding_options_alloc()
        if bilinear:
            # color_conversion_options sits behind decoder_id (offset 56): version u8, downsampling enum @60, upsampling enum @64,
            # only_use_preferred_chroma_algorithm u8 @68 — what heif-dec -C bilinear sets (examples/heif_dec.cc:502-509);
            # heif_chroma_upsampling_bilinear = 2 (heif.h)
            C.memmove(opts + 64, C.byref(C.c_int(2)), 4)
            C.memmove(opts + 68, C.byref(C.c_uint8(1)), 1)
        if decoder_id is not None:
            # struct heif_decoding_options (heif.h:1565-1611): decoder_id is the const char* after
            # version(u8) ignore_transformations(u8) start_progress/on_progress/end_progress/progress_user_data
            # (4 pointers) convert_hdr_to_8bit(u8) strict_decoding(u8) -> offset 8+32+8 = 48
            did = C.c_char_p(decoder_id.encode())
            C.memmove(opts + 48, C.byref(did), 8)
        try:
            _check(L.heif_decode_image(h, C.byref(img), colorspace, chroma, opts), "decode")
        finally:
            if opts:
                L.heif_decoding_options_free(opts)
        if premultiply:      # heif_image_rgba_premultiply_alpha (heif.cc:1444): interleaved RGBA only
            L.heif_image_rgba_premultiply_alpha.restype = HeifError
            L.heif_image_rgba_premultiply_alpha.argtypes = [C.c_void_p]
            _check(L.heif_image_rgba_premultiply_alpha(img), "premultiply")
        out = {"colorspace": L.heif_image_get_colorspace(img), "chroma": L.heif_image_get_chroma_format(img)}
        if chroma in (CHROMA_RGB, CHROMA_RGBA):
            bpp = 3 if chroma == CHROMA_RGB else 4
            out["interleaved"] = plane_bytes(img, CH_INTERLEAVED, bpp)
        elif chroma in (CHROMA_RRGGBB_BE, CHROMA_RRGGBB_LE, CHROMA_RRGGBBAA_BE, CHROMA_RRGGBBAA_LE):
            bpp = 6 if chroma in (CHROMA_RRGGBB_BE, CHROMA_RRGGBB_LE) else 8
            out["interleaved"] = plane_bytes(img, CH_INTERLEAVED, bpp)
            out["bpp"] = L.heif_image_get_bits_per_pixel_range(img, CH_INTERLEAVED)
        else:
            for name, ch in (("Y", CH_Y), ("Cb", CH_CB), ("Cr", CH_CR), ("A", CH_ALPHA)):
                if L.heif_image_has_channel(img, ch):
                    bits = L.heif_image_get_bits_per_pixel_range(img, ch)
                    out[name] = plane_bytes(img, ch, (bits + 7) // 8)
                    out["bpp"] = bits
        L.heif_image_release(img)
        L.heif_image_handle_release(h)
        return out
    finally:
        L.heif_context_free(ctx)


def top_level_ids(data):
    L = lib()
    ctx = L.heif_context_alloc()
    buf = C.create_string_buffer(data, len(data))
    _check(L.heif_context_read_from_memory_without_copy(ctx, buf, len(data), None), "read")
    n = L.heif_context_get_number_of_top_level_images(ctx)
    ids = (C.c_uint32 * n)()
    L.heif_context_get_list_of_top_level_image_IDs(ctx, ids, n)
    L.heif_context_free(ctx)
    return list(ids)

"""Host side of the hot path, mirroring the reference's interface (src/decoder_host.cpp) over the C ABI.

    reference                                              here
    -----------------------------------------------------  ---------------------------------------------------
    DpuSet::allocate + load            (:32, :268)         Decoder(device)
    pim.copy / pim.exec / pim.copy     (:276-308)          Decoder.exec_mcus(metadata, mcus)      [compat layout]
    read_JPEG                          (jpeg_scanner:345)  parse_header(bytes)
    decode_Huffman_data + exec + write_BMP gather          Decoder.decode(files, fmt) / Batch     [fast layout]
    main(argv): sort by size, decode, write <name>.bmp     decode_files(paths)
"""
import ctypes as C
import os

import numpy as np

from . import _lib as L


def parse_header(data):
    """read_JPEG on in-memory bytes (minus the scan copy).  Returns (status, ImageDesc)."""
    d = L.ImageDesc()
    buf = np.frombuffer(data, dtype=np.uint8)
    st = L.lib().bj_parse_header(buf.ctypes.data_as(C.c_void_p), len(data), C.byref(d))
    return st, d


class PinnedBuffer:
    """Page-locked host memory from bj_host_alloc, exposed as a numpy uint8 array."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.ptr = L.lib().bj_host_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise MemoryError(f"bj_host_alloc({nbytes}) failed")
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr))[: self.nbytes]

    def free(self):
        if self.ptr:
            self.array = None
            L.lib().bj_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _file_arrays(files):
    """list of bytes / uint8 arrays -> (keepalive arrays, char** , size_t*)"""
    arrs = [f if isinstance(f, np.ndarray) else np.frombuffer(f, dtype=np.uint8) for f in files]
    n = len(arrs)
    ptrs = (C.c_void_p * max(n, 1))(*[a.__array_interface__["data"][0] for a in arrs])
    lens = (C.c_size_t * max(n, 1))(*[a.size for a in arrs])
    return arrs, ptrs, lens


class Batch:
    """A device-resident batch (bj_batch): upload -> decode -> download, each asynchronous on the context's stream."""

    def __init__(self, dec, files, fmt=L.BJ_OUT_RGB8):
        self.dec = dec
        self.n = len(files)
        self._keep, ptrs, lens = _file_arrays(files)
        self.h = C.c_void_p()
        L.check(L.lib().bj_batch_create(dec.ctx, ptrs, lens, self.n, fmt, C.byref(self.h)), "bj_batch_create", dec.ctx)
        self.offsets, self.sizes, self.parse_status = [], [], []
        for i in range(self.n):
            off, nb = C.c_size_t(), C.c_size_t()
            st = L.lib().bj_batch_output_offset(self.h, i, C.byref(off), C.byref(nb))
            self.parse_status.append(st)
            self.offsets.append(off.value)
            self.sizes.append(nb.value)
        self.total_out = max((o + s for o, s in zip(self.offsets, self.sizes)), default=0)

    def upload(self, stream=None):
        """stream: a cudaStream_t as an integer (e.g. torch.cuda.Stream().cuda_stream); None = the context's own."""
        L.check(L.lib().bj_batch_upload(self.h, stream), "bj_batch_upload", self.dec.ctx)

    def decode(self, stream=None):
        L.check(L.lib().bj_batch_decode(self.h, stream), "bj_batch_decode", self.dec.ctx)

    def sync(self):
        L.check(L.lib().bj_batch_sync(self.h), "bj_batch_sync", self.dec.ctx)

    def download(self, out=None, only=None):
        """Copy decoded images to the host.  `out`: uint8 array of total_out bytes (e.g. PinnedBuffer.array) laid
        out like the device buffer; `only`: indices to copy (default all).  Returns per-image views (None = skipped)."""
        if out is None:
            out = np.zeros(self.total_out, dtype=np.uint8)
        base = out.ctypes.data
        want = set(range(self.n)) if only is None else set(only)
        ok = [st == L.BJ_OK and i in want for i, st in enumerate(self.parse_status)]
        ptrs = (C.c_void_p * max(self.n, 1))(*[base + o if k else None for o, k in zip(self.offsets, ok)])
        L.check(L.lib().bj_batch_download(self.h, ptrs, None), "bj_batch_download", self.dec.ctx)
        return [out[o:o + s] if k else None for o, s, k in zip(self.offsets, self.sizes, ok)]

    def status(self):
        st = (C.c_int * max(self.n, 1))()
        L.check(L.lib().bj_batch_status(self.h, st), "bj_batch_status")
        return list(st)[: self.n]

    def info(self):
        info = L.BatchInfo()
        L.check(L.lib().bj_batch_get_info(self.h, C.byref(info)), "bj_batch_get_info")
        return info

    def device_output(self, i):
        p, nb = C.c_void_p(), C.c_size_t()
        L.check(L.lib().bj_batch_device_output(self.h, i, C.byref(p), C.byref(nb)), "bj_batch_device_output")
        return p.value, nb.value

    def destroy(self):
        if self.h:
            L.lib().bj_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Job:
    """A batch handed to bj_submit: wait() blocks until its outputs are in host memory and returns the status array."""

    def __init__(self, dec, handle, keep, status, n):
        self.dec, self.h, self._keep, self._status, self.n = dec, handle, keep, status, n

    def wait(self):
        if self.h is not None:
            h, self.h = self.h, None
            L.check(L.lib().bj_wait(h), "bj_wait", self.dec.ctx)
        return self._status[: self.n]


class Decoder:
    """One context.  Decoder(device) drives one GPU (one process per GPU, the torchrun way); Decoder(devices=[...])
    - or devices="all" - drives several GPUs of the box from this one process (bj_create_multi), the way the
    reference's single process takes every DPU (src/decoder_host.cpp:32-33)."""

    def __init__(self, device=0, devices=None):
        self.ctx = C.c_void_p()
        if devices is None:
            st = L.lib().bj_create(C.byref(self.ctx), device)
            what = "bj_create"
        elif isinstance(devices, str):
            st = L.lib().bj_create_multi(C.byref(self.ctx), None, 0)
            what = "bj_create_multi"
        else:
            arr = (C.c_int * len(devices))(*devices)
            st = L.lib().bj_create_multi(C.byref(self.ctx), arr, len(devices))
            what = "bj_create_multi"
        if st != L.BJ_OK:
            self.ctx = None
            raise L.BjError(st, what + " (this back end has no CPU fallback)")
        self.device = device
        self.device_count = L.lib().bj_device_count(self.ctx)

    def close(self):
        if self.ctx:
            L.lib().bj_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name, value):
        L.check(L.lib().bj_set_option(self.ctx, name.encode(), int(value)), f"bj_set_option({name})")

    def stat(self, name):
        v = C.c_double()
        L.check(L.lib().bj_get_stat(self.ctx, name.encode(), C.byref(v)), f"bj_get_stat({name})")
        return v.value

    # ---- compat entry: the DPU program on the reference's own buffers (in place)
    def exec_mcus(self, metadata, mcus):
        md = np.ascontiguousarray(metadata, dtype=np.uint32).reshape(-1, 276)
        out = np.ascontiguousarray(mcus, dtype=np.int16).copy()
        L.check(L.lib().bj_exec_mcus(self.ctx, md.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), md.shape[0]),
                "bj_exec_mcus", self.ctx)
        return out

    # ---- stage entries (known-answer tests)
    def stage_entropy(self, data):
        st, d = parse_header(data)
        if st != L.BJ_OK:
            raise L.BjError(st, "bj_parse_header")
        nmx = (d.mcu_w + d.hs - 1) // d.hs
        nmy = (d.mcu_h + d.vs - 1) // d.vs
        bpm = sum(d.comp_h[j] * d.comp_v[j] for j in range(d.ncomp))
        coef = np.full((nmx * nmy * bpm, 64), 0x5A5A, dtype=np.int16)
        buf = np.frombuffer(data, dtype=np.uint8)
        status = C.c_int(0)
        L.check(L.lib().bj_stage_entropy(self.ctx, buf.ctypes.data_as(C.c_void_p), len(data), coef.ctypes.data_as(C.c_void_p),
                                         coef.nbytes, C.byref(status)), "bj_stage_entropy", self.ctx)
        return coef, status.value

    def stage_idct_color(self, desc, coef_zz, fmt=L.BJ_OUT_RGB8):
        co = np.ascontiguousarray(coef_zz, dtype=np.int16)
        out = np.zeros(L.lib().bj_output_size(C.byref(desc), fmt), dtype=np.uint8)
        L.check(L.lib().bj_stage_idct_color(self.ctx, C.byref(desc), co.ctypes.data_as(C.c_void_p), fmt, out.ctypes.data_as(C.c_void_p)),
                "bj_stage_idct_color", self.ctx)
        return out

    # ---- full path, one call (host buffers in, host buffers out)
    def decode(self, files, fmt=L.BJ_OUT_RGB8, outs=None):
        """Returns (list of uint8 arrays or None for rejected files, list of per-image status)."""
        n = len(files)
        keep, ptrs, lens = _file_arrays(files)
        if outs is None:
            outs = []
            for f in files:
                st, d = parse_header(f)
                outs.append(np.zeros(L.lib().bj_output_size(C.byref(d), fmt), dtype=np.uint8) if st == L.BJ_OK else None)
        optrs = (C.c_void_p * max(n, 1))(*[o.__array_interface__["data"][0] if o is not None else None for o in outs])
        status = (C.c_int * max(n, 1))()
        L.check(L.lib().bj_decode_batch(self.ctx, ptrs, lens, n, fmt, optrs, status), "bj_decode_batch", self.ctx)
        st = list(status)[:n]
        return [o if s in (L.BJ_OK, L.BJ_ERR_CORRUPT_SCAN) else None for o, s in zip(outs, st)], st


    def decode_desc(self, descs, scans, kinds=None, fmt=L.BJ_OUT_REF_MCUS):
        """bj_decode_batch_desc: the caller holds parsed headers (ImageDesc, e.g. filled from the reference's `Header`) and
        the scan bytes.  Returns (list of arrays - int16 for BJ_OUT_REF_MCUS, uint8 otherwise -, list of status)."""
        n = len(descs)
        arr = (L.ImageDesc * max(n, 1))(*descs)
        keep, ptrs, lens = _file_arrays(scans)
        m = C.c_int(0)
        sizes = [L.lib().bj_ref_mcus_size(C.byref(d), 100, C.byref(m)) if fmt == L.BJ_OUT_REF_MCUS else L.lib().bj_output_size(C.byref(d), fmt) for d in descs]
        outs = [np.zeros(s, dtype=np.uint8) for s in sizes]
        optrs = (C.c_void_p * max(n, 1))(*[o.__array_interface__["data"][0] for o in outs])
        kp = (C.c_int * max(n, 1))(*kinds) if kinds is not None else None
        status = (C.c_int * max(n, 1))()
        L.check(L.lib().bj_decode_batch_desc(self.ctx, arr, ptrs, lens, kp, n, fmt, optrs, status), "bj_decode_batch_desc", self.ctx)
        if fmt == L.BJ_OUT_REF_MCUS:
            outs = [o.view(np.int16) for o in outs]
        return outs, list(status)[:n]

    def decode_packed(self, src, src_off, src_len, dst, dst_off, fmt=L.BJ_OUT_RGB8):
        """bj_decode_batch for a batch-pipeline caller: the n files sit in ONE host buffer `src` (uint8 array, ideally
        a PinnedBuffer) at byte offsets `src_off` with lengths `src_len`; image i is written to `dst[dst_off[i]:]`
        (bj_output_size bytes).  The pointer tables are built with vector arithmetic, so the host-side cost per
        call does not grow with a Python loop over images.  Returns the per-image status as an int32 array."""
        n = len(src_off)
        ip = (np.asarray(src_off, dtype=np.uint64) + np.uint64(src.__array_interface__["data"][0]))
        il = np.ascontiguousarray(src_len, dtype=np.uint64)
        op = (np.asarray(dst_off, dtype=np.uint64) + np.uint64(dst.__array_interface__["data"][0]))
        status = np.zeros(max(n, 1), dtype=np.int32)
        L.check(L.lib().bj_decode_batch(self.ctx, ip.ctypes.data_as(C.c_void_p), il.ctypes.data_as(C.c_void_p), n, fmt,
                                        op.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p)), "bj_decode_batch", self.ctx)
        return status[:n]


    def submit_packed(self, src, src_off, src_len, dst, dst_off, fmt=L.BJ_OUT_RGB8):
        """decode_packed without waiting (bj_submit): returns a Job; job.wait() gives the status array."""
        n = len(src_off)
        ip = (np.asarray(src_off, dtype=np.uint64) + np.uint64(src.__array_interface__["data"][0]))
        il = np.ascontiguousarray(src_len, dtype=np.uint64)
        op = (np.asarray(dst_off, dtype=np.uint64) + np.uint64(dst.__array_interface__["data"][0]))
        status = np.zeros(max(n, 1), dtype=np.int32)
        h = C.c_void_p()
        L.check(L.lib().bj_submit(self.ctx, ip.ctypes.data_as(C.c_void_p), il.ctypes.data_as(C.c_void_p), n, fmt,
                                  op.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p), C.byref(h)), "bj_submit", self.ctx)
        return Job(self, h, (ip, il, op, src, dst), status, n)


def lpt_shards(costs, world_size):
    """Deal a stream of images over `world_size` ranks: longest processing time first - sort by cost (compressed size,
    the reference's sort key, src/decoder_host.cpp:46-61) descending and give each image to the rank with the least
    work so far.  Returns world_size index lists (each in ascending index order)."""
    import heapq
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    heap = [(0, r) for r in range(world_size)]
    out = [[] for _ in range(world_size)]
    for i in order:
        load, r = heapq.heappop(heap)
        out[r].append(i)
        heapq.heappush(heap, (load + costs[i], r))
    return [sorted(o) for o in out]


def shard_by_size(sizes, world_size):
    """Deal images to ranks: sort ascending by compressed size (the reference sorts its inputs the same way,
    src/decoder_host.cpp:46-61,360), then round-robin.  Images are independent, so there is no collective on the
    data path.  Returns world_size index lists."""
    order = sorted(range(len(sizes)), key=lambda i: (sizes[i], i))
    return [order[r::world_size] for r in range(world_size)]


def decode_files(paths, device=0, rank=0, world_size=1, decoder=None):
    """The reference CLI's behaviour (src/decoder_host.cpp:352-394): every input gets `<name>.bmp` next to it
    (:328-330); unreadable / invalid files are reported and skipped (:120-123).  With world_size > 1 each rank
    takes its shard.  Returns {path: status}."""
    paths = list(paths)
    sizes = [os.path.getsize(p) if os.path.exists(p) else 0 for p in paths]
    mine = shard_by_size(sizes, world_size)[rank]
    dec = decoder or Decoder(device)
    files = []
    for i in mine:
        try:
            with open(paths[i], "rb") as f:
                files.append(f.read())
        except OSError:
            files.append(b"")
    outs, status = dec.decode(files, L.BJ_OUT_BMP)
    result = {}
    for i, o, st in zip(mine, outs, status):
        result[paths[i]] = st
        if o is None:
            print(f"Error - {paths[i]}: {L.lib().bj_status_string(st).decode()}")
            continue
        stem, dot, _ = paths[i].rpartition(".")
        with open((stem if dot else paths[i]) + ".bmp", "wb") as f:
            f.write(o.tobytes())
    return result

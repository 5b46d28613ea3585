"""H2D bandwidth of one step's features (115 MB, in 5-image pieces like sqd_head_detect_host) from ordinary pinned host memory
and from write-combined pinned memory (cudaHostAllocWriteCombined).  usage: python tools/h2d_wc.py"""
import ctypes as C
import torch

torch.cuda.init()
rt = C.CDLL("libcudart.so.12")
n = 20 * 768 * 24 * 78 * 4
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream()


def alloc(flags):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flags)) == 0
    C.memset(p, 1, n)
    return p


def bw(p, pieces=4, reps=30):
    rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    piece = n // pieces
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for r in range(reps + 3):
        if r == 3:
            e0.record(st)
        for i in range(pieces):
            assert rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr() + i * piece), C.c_void_p(p.value + i * piece), piece, 1,
                                      C.c_void_p(st.cuda_stream)) == 0
    e1.record(st)
    torch.cuda.synchronize()
    return n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


for name, flags in (("pinned (default)", 0), ("pinned + portable", 1), ("write-combined", 4), ("write-combined + mapped", 4 | 2)):
    p = alloc(flags)
    print(f"{name:26s} {bw(p):6.2f} GB/s")
    rt.cudaFreeHost(p)

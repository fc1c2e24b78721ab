"""Host -> device bandwidth of the copies the host-streamed path issues: pinned rvs array in the reference's layout
(particle-major rows of n_obs doubles), strided 2-D copies of `chunk` time steps per row against one contiguous
copy, on one and on two copy streams.  usage: probe_h2d.py [logN]"""
import sys
import time

import torch
from cuda.bindings import runtime as rt

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n, nobs = 1 << logn, 1001
dev = torch.device("cuda:0")
host = torch.empty(n * nobs, dtype=torch.float64).pin_memory()
host.normal_()
stage = torch.empty(n * 1024, dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
H2D = rt.cudaMemcpyKind.cudaMemcpyHostToDevice


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nbytes = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return nbytes / best / 1e9


def contiguous(nbytes):
    def f():
        (err,) = rt.cudaMemcpyAsync(stage.data_ptr(), host.data_ptr(), nbytes, H2D, s1.cuda_stream)
        assert err == rt.cudaError_t.cudaSuccess, err
        return nbytes
    return f


def strided(chunk, streams, chunks=3, split_rows=1):
    def f():
        tot = 0
        for c in range(chunks):
            for h in range(split_rows):
                st = streams[(c * split_rows + h) % len(streams)]
                rows = n // split_rows
                src = host.data_ptr() + (c * chunk + h * rows * nobs) * 8
                dst = stage.data_ptr() + (c % 2) * n * chunk * 8 + h * rows * chunk * 8
                (err,) = rt.cudaMemcpy2DAsync(dst, chunk * 8, src, nobs * 8, chunk * 8, rows, H2D, st.cuda_stream)
                assert err == rt.cudaError_t.cudaSuccess, err
                tot += chunk * 8 * rows
        return tot
    return f


print("contiguous 2 GiB               %.1f GB/s" % timed(contiguous(1 << 31)))
for chunk in (64, 128, 256, 512):
    print("2-D rows of %4d B, 1 stream    %.1f GB/s" % (chunk * 8, timed(strided(chunk, [s1]))))
print("2-D rows of 2048 B, 2 streams (alternating chunks)  %.1f GB/s" % timed(strided(256, [s1, s2], chunks=4)))
print("2-D rows of 2048 B, 2 streams (rows split in two)   %.1f GB/s" % timed(strided(256, [s1, s2], chunks=3, split_rows=2)))
print("2-D rows of 4096 B, 2 streams (rows split in two)   %.1f GB/s" % timed(strided(512, [s1, s2], chunks=1, split_rows=2)))
print("2-D rows of 2048 B, 4 copies in flight on 2 streams (rows split in four) %.1f GB/s" % timed(strided(256, [s1, s2], chunks=3, split_rows=4)))

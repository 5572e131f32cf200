"""Why does posix_fallocate of ~20 pieces of an already allocated tmpfs file take 100+ ms per rank?  (host-only probe)"""
import mmap, os, sys, threading, time
from multiprocessing import Process, Barrier

path = "/dev/shm/falloc_probe.bin"
TOTAL = 420 << 20


def worker(rank, world, bar, mode):
    fd = os.open(path, os.O_RDWR | os.O_CREAT, 0o644)
    stop = False
    th = None
    if rank == 0 and mode != "noprealloc":
        def pre():
            off = 0
            while off < (2100 << 20) and not stop:
                os.posix_fallocate(fd, off, 64 << 20)
                off += 64 << 20
        th = threading.Thread(target=pre)
        th.start()
    time.sleep(0.045)  # the pile-up
    bar.wait()
    t0 = time.perf_counter()
    if th:
        stop = True
        th.join()
    t1 = time.perf_counter()
    if os.fstat(fd).st_size != TOTAL:
        os.ftruncate(fd, TOTAL)
    t2 = time.perf_counter()
    n = 23
    piece = TOTAL // (n * world)
    times = []
    for k in range(n):
        a0 = (k * world + rank) * piece + 1234
        t = time.perf_counter()
        if mode == "fallocate_syscall":
            import ctypes
            libc = ctypes.CDLL(None, use_errno=True)
            libc.fallocate(fd, 0, ctypes.c_long(a0), ctypes.c_long(piece - 2000))
        else:
            os.posix_fallocate(fd, a0, piece - 2000)
        times.append(1e3 * (time.perf_counter() - t))
    t3 = time.perf_counter()
    print(f"[{mode}] rank {rank}: join {1e3*(t1-t0):.1f} ms, size {1e3*(t2-t1):.1f} ms, {n} fallocates {1e3*(t3-t2):.1f} ms (max {max(times):.1f}, first {times[0]:.1f})", flush=True)
    os.close(fd)


for mode in ("prealloc", "noprealloc", "fallocate_syscall"):
    for rep in range(2):
        if os.path.exists(path):
            os.unlink(path)
        bar = Barrier(2)
        ps = [Process(target=worker, args=(r, 2, bar, mode)) for r in range(2)]
        [p.start() for p in ps]
        [p.join() for p in ps]
os.unlink(path)

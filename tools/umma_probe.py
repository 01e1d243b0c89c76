"""tcgen05 probe driver (needs a B200): pins down the descriptor conventions, TMEM layouts, accumulator
rounding and issue rate of tcgen05.mma kind::tf32 through pinn_umma_probe.  Each group runs in its
own process (a bad descriptor kills the CUDA context).  `python tools/umma_probe.py` runs all groups."""
import ctypes as C
import subprocess
import sys

import numpy as np

sys.path.insert(0, ".")


def _lib():
    from pinn_based_online_pde_calculator_b200.engine import load_library

    lib = load_library()
    lib.pinn_umma_probe.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double),
                                    C.c_void_p]
    return lib


def raw(lib, A_img, B_img, B1_img, M, N, ksteps, a_mn, b_mn, a_desc, b_desc, nsets=1, reps=1, nd=1, a_lt=0, b_lt=0, a_tmem=0):
    """a_desc/b_desc = (lbo, sbo, step) in bytes"""
    A_img = np.ascontiguousarray(A_img, np.float32)
    B_img = np.ascontiguousarray(B_img, np.float32)
    B1c = np.ascontiguousarray(B1_img, np.float32) if B1_img is not None else None
    cfg = np.array([M, N, ksteps, a_mn, b_mn, nsets, reps, A_img.size, B_img.size, *a_desc, *b_desc, nd, a_lt, b_lt, a_tmem], np.int32)
    out = np.empty((128, 512), np.float32)
    cyc = C.c_double()
    st = np.zeros(2, np.int32)
    rc = lib.pinn_umma_probe(0, A_img.ctypes.data, B_img.ctypes.data, B1c.ctypes.data if B1c is not None else None,
                             cfg.ctypes.data, out.ctypes.data, C.byref(cyc), st.ctypes.data)
    return rc, out, cyc.value, st


def img_k_major(X):
    """X[R][K] -> un-swizzled K-major image; (lbo, sbo, step) = (R*16, 128, 2*R*16)"""
    R, K = X.shape
    img = np.zeros(R * K, np.float32)
    r, k = np.meshgrid(np.arange(R), np.arange(K), indexing="ij")
    off = (k // 4) * (R * 4) + r * 4 + (k % 4)
    img[off.ravel()] = X.ravel()
    return img, (R * 16, 128, 2 * R * 16)


def img_mn_major(X, variant=0):
    """X[R][K] (R = M/N extent) -> un-swizzled MN-major image: 4 consecutive mn per 16 B, 8 k rows per core matrix"""
    R, K = X.shape
    img = np.zeros(R * K, np.float32)
    r, k = np.meshgrid(np.arange(R), np.arange(K), indexing="ij")
    off = (r // 4) * (K * 4) + (k // 8) * 32 + (k % 8) * 4 + (r % 4)
    img[off.ravel()] = X.ravel()
    return img, ((128, K * 16, 128) if variant == 0 else (K * 16, 128, 128))


def lanes_m64(s):
    r = np.arange(64)
    return 32 * (r // 16) + 16 * s + (r % 16)


def run_k(lib, A, B, B1=None, nsets=1, reps=1, nd=1):
    M, K = A.shape
    ai, ad = img_k_major(A)
    bi, bd = img_k_major(B)
    b1 = img_k_major(B1)[0] if B1 is not None else None
    return raw(lib, ai, bi, b1, M, B.shape[0], K // 8, 0, 0, ad, bd, nsets, reps, nd)


def group_kmajor():
    lib = _lib()
    rng = np.random.RandomState(0)
    ints = lambda *s: rng.randint(-4, 5, size=s).astype(np.float32)
    for (M, N, K) in [(128, 64, 64), (128, 256, 32), (128, 8, 8), (64, 64, 64), (64, 32, 16)]:
        A, B = ints(M, K), ints(N, K)
        rc, out, cyc, st = run_k(lib, A, B)
        ref = A @ B.T
        got = out[:M, :N] if M == 128 else out[lanes_m64(0), :N]
        print(f"K-major M={M} N={N} K={K}: rc={rc} done={st[0]} match={np.array_equal(got, ref)} cycles={cyc:.0f}")
        if M == 64 and not np.array_equal(got, ref):
            for r in (0, 1, 15, 16, 17, 32, 48, 63):
                print(f"   row {r} found in lanes {[l for l in range(128) if np.array_equal(out[l, :N], ref[r])]}")
    A, B0, B1 = ints(64, 64), ints(64, 64), ints(64, 64)
    rc, out, cyc, st = run_k(lib, A, B0, B1, nsets=2)
    print("two interleaved M=64 tiles: set0", np.array_equal(out[lanes_m64(0), :64], A @ B0.T), "set1",
          np.array_equal(out[lanes_m64(1), :64], A @ B1.T), "done", st[0])


def group_mn_discover():
    """A image holds its own word index; B = identity (K-major) -> D[m][k] = word index the hardware read for A(m,k)."""
    lib = _lib()
    words = 45056  # 176 KB
    w = np.arange(words)
    eye, bd = img_k_major(np.eye(8, dtype=np.float32))
    for (lbo, sbo) in [(128, 256), (256, 128), (16, 32), (1024, 512)]:
        res = []
        for part in (w % 2048, w // 2048):
            rc, out, cyc, st = raw(lib, part.astype(np.float32), eye, None, 128, 8, 1, 1, 0, (lbo, sbo, 0), bd)
            res.append(out[:, :8].copy())
        W = (res[0] + 2048 * res[1]).astype(np.int64)  # [m][k] -> word offset
        print(f"A MN-major discovery, LBO={lbo} SBO={sbo}: rc={rc} done={st[0]}")
        print("  byte offsets of (m=0..11, k=0):", (4 * W[:12, 0]).tolist())
        print("  byte offsets of (m=0, k=0..7):", (4 * W[0, :8]).tolist())
        print("  byte offsets of (m=0,8,16,..,120, k=0):", (4 * W[0:128:8, 0]).tolist())
        print("  byte offsets of (m=1, k=0..7):", (4 * W[1, :8]).tolist())
    # K-major control with the same machinery (must reproduce the K-major formula)
    res = []
    for part in (w % 2048, w // 2048):
        rc, out, cyc, st = raw(lib, part.astype(np.float32), eye, None, 128, 8, 1, 0, 0, (2048, 128, 0), bd)
        res.append(out[:, :8].copy())
    W = (res[0] + 2048 * res[1]).astype(np.int64)
    print("K-major control LBO=2048 SBO=128: (m=0..9,k=0):", (4 * W[:10, 0]).tolist(), " (m=0,k=0..7):", (4 * W[0, :8]).tolist())
    # B operand MN-major: A = delta K-major, D[m][n] = B(n, k=m) for m < 8
    A = np.zeros((128, 8), np.float32)
    A[:8, :8] = np.eye(8)
    ai, ad = img_k_major(A)
    for (lbo, sbo) in [(128, 256), (256, 128)]:
        res = []
        for part in (w % 2048, w // 2048):
            rc, out, cyc, st = raw(lib, ai, part.astype(np.float32)[:20000], None, 128, 64, 1, 0, 1, ad, (lbo, sbo, 0))
            res.append(out[:8, :64].copy())
        W = (res[0] + 2048 * res[1]).astype(np.int64).T  # [n][k]
        print(f"B MN-major discovery, LBO={lbo} SBO={sbo}: rc={rc} done={st[0]}")
        print("  byte offsets of (n=0..11, k=0):", (4 * W[:12, 0]).tolist())
        print("  byte offsets of (n=0, k=0..7):", (4 * W[0, :8]).tolist())
        print("  byte offsets of (n=0,8,..,56, k=0):", (4 * W[0:64:8, 0]).tolist())


def group_mn_check():
    lib = _lib()
    rng = np.random.RandomState(1)
    ints = lambda *s: rng.randint(-4, 5, size=s).astype(np.float32)
    for variant in (0, 1):
        for (M, N, K) in [(128, 64, 64), (64, 64, 32)]:
            A, B = ints(M, K), ints(N, K)
            ref = A @ B.T
            for (amn, bmn) in [(1, 0), (0, 1), (1, 1)]:
                ai, ad = img_mn_major(A, variant) if amn else img_k_major(A)
                bi, bd = img_mn_major(B, variant) if bmn else img_k_major(B)
                rc, out, cyc, st = raw(lib, ai, bi, None, M, N, K // 8, amn, bmn, ad, bd)
                got = out[:M, :N] if M == 128 else out[lanes_m64(0), :N]
                print(f"MN check variant {variant} M={M} N={N} K={K} A_mn={amn} B_mn={bmn}: done={st[0]} match={np.array_equal(got, ref)}")


def group_rounding():
    lib = _lib()
    for sign in (1.0, -1.0):
        M, N, K = 128, 8, 128
        A = np.zeros((M, K), np.float32)
        B = np.zeros((N, K), np.float32)
        A[:, 0] = sign
        B[:, 0] = 1.0
        for j in range(1, K // 8):
            A[:, 8 * j] = sign * 2.0 ** -12
            B[:, 8 * j] = 1.5 * 2.0 ** -12
        rc, out, cyc, st = run_k(lib, A, B)
        print(f"accumulate across MMAs, sign {sign:+.0f}: 15 adds of 0.75 ulp -> (D - sign)/ulp = {(out[0, 0] - sign) / 2.0 ** -23:+.2f}"
              f"  (RN +-15, truncate 0, exact +-11.25; done={st[0]})")
    A = np.zeros((128, 8), np.float32); B = np.zeros((8, 8), np.float32)
    A[:, 0] = 1.0; B[:, 0] = 1.0
    A[:, 1:] = 2.0 ** -12; B[:, 1:] = 1.5 * 2.0 ** -12
    rc, out, cyc, st = run_k(lib, A, B)
    print(f"inside one MMA: 1 + 7 x 0.75 ulp -> (D-1)/ulp = {(out[0, 0] - 1) / 2.0 ** -23:+.2f} (exact 5.25)")
    A = np.zeros((128, 8), np.float32); B = np.zeros((8, 8), np.float32)
    A[:, 0] = 1.0 + 2.0 ** -11 + 2.0 ** -12; B[:, 0] = 1.0
    rc, out, cyc, st = run_k(lib, A, B)
    print(f"input conversion of 1+2^-11+2^-12: (D-1)*2^11 = {(out[0, 0] - 1) * 2 ** 11:.3f} (1.0 = truncated to tf32, 2.0 = rounded, 1.5 = all bits used)")
    # long chain with random data: compare against float64
    rng = np.random.RandomState(3)
    for K in (64, 192):
        A = rng.standard_normal((128, K)).astype(np.float32)
        B = rng.standard_normal((64, K)).astype(np.float32)
        tf = lambda x: (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
        A, B = tf(A), tf(B)
        rc, out, cyc, st = run_k(lib, A, B)
        ref = A.astype(np.float64) @ B.astype(np.float64).T
        err = out[:128, :64] - ref
        print(f"random tf32-exact inputs K={K}: mean err/|ref|rms = {err.mean() / np.sqrt((ref ** 2).mean()):+.3e}, rms = "
              f"{np.sqrt((err ** 2).mean()) / np.sqrt((ref ** 2).mean()):.3e}, signed bias vs sign(ref) = "
              f"{(err * np.sign(ref)).mean() / np.abs(ref).mean():+.3e}")


def group_rate():
    lib = _lib()
    rng = np.random.RandomState(0)
    ints = lambda *s: rng.randint(-4, 5, size=s).astype(np.float32)
    for (M, N) in [(128, 64), (128, 128), (128, 256), (64, 64), (64, 32), (64, 128), (64, 256), (128, 32)]:
        K = 64
        A, B = ints(M, K), ints(N, K)
        for nd in (1, 2):
            if nd * N > 512:
                continue
            t = {}
            for reps in (16, 64):
                rc, out, cyc, st = run_k(lib, A, B, reps=reps, nd=nd)
                t[reps] = cyc
            n = (64 - 16) * K // 8
            print(f"M={M} N={N} accumulators={nd}: {(t[64] - t[16]) / n:.1f} cycles per tcgen05.mma (K=8)  -> "
                  f"{2 * M * N * 8 / ((t[64] - t[16]) / n):.0f} flop/cycle/SM")
    A, B0, B1 = ints(64, 64), ints(256, 64), ints(256, 64)
    t = {}
    for reps in (16, 64):
        rc, out, cyc, st = run_k(lib, A, B0, B1, nsets=2, reps=reps)
        t[reps] = cyc
    n = (64 - 16) * 2 * 8
    print(f"two interleaved M=64 N=256 products: {(t[64] - t[16]) / n:.1f} cycles per tcgen05.mma")


def group_ts():
    """A operand from TMEM (lane = row, 8 columns per k-step), B K-major in shared memory"""
    lib = _lib()
    rng = np.random.RandomState(2)
    ints = lambda *s: rng.randint(-4, 5, size=s).astype(np.float32)
    for (N, K) in [(64, 64), (256, 32), (8, 8)]:
        A, B = ints(128, K), ints(N, K)
        bi, bd = img_k_major(B)
        rc, out, cyc, st = raw(lib, A, bi, None, 128, N, K // 8, 0, 0, (0, 0, 0), bd, a_tmem=1)
        print(f"TS mode M=128 N={N} K={K}: rc={rc} done={st[0]} match={np.array_equal(out[:128, :N], A @ B.T)}")
    for N in (64, 128, 256):
        A, B = ints(128, 64), ints(N, 64)
        bi, bd = img_k_major(B)
        t = {}
        for reps in (16, 64):
            rc, out, cyc, st = raw(lib, A, bi, None, 128, N, 8, 0, 0, (0, 0, 0), bd, reps=reps, a_tmem=1)
            t[reps] = cyc
        print(f"TS mode M=128 N={N}: {(t[64] - t[16]) / (48 * 8):.1f} cycles per tcgen05.mma")


def group_mn_swizzle():
    """MN-major A with the swizzled layout types: which byte does the hardware read for A(m, k)?"""
    lib = _lib()
    words = 45056
    w = np.arange(words)
    eye, bd = img_k_major(np.eye(8, dtype=np.float32))
    for lt in (2, 4, 6, 1):
        for (lbo, sbo) in [(4096, 1024), (1024, 4096)]:
            res = []
            for part in (w % 2048, w // 2048):
                rc, out, cyc, st = raw(lib, part.astype(np.float32), eye, None, 128, 8, 1, 1, 0, (lbo, sbo, 0), bd, a_lt=lt)
                res.append(out[:, :8].copy())
            W = 4 * (res[0] + 2048 * res[1]).astype(np.int64)
            print(f"A MN-major layout_type={lt} LBO={lbo} SBO={sbo}: rc={rc} done={st[0]}")
            print("  bytes (m=0..7, k=0):", W[:8, 0].tolist())
            print("  bytes (m=0,4,..,60, k=0):", W[0:64:4, 0].tolist())
            print("  bytes (m=64,96,  k=0):", W[[64, 96], 0].tolist())
            for k in range(8):
                print(f"  bytes (m=0,4,..,28, k={k}):", W[0:32:4, k].tolist())
    # K-major with 128B swizzle (the layout TMA would write): row r, 32 k-elements per 128 B row
    for lt in (2,):
        res = []
        for part in (w % 2048, w // 2048):
            rc, out, cyc, st = raw(lib, part.astype(np.float32), eye, None, 128, 8, 1, 0, 0, (16, 1024, 0), bd, a_lt=lt)
            res.append(out[:, :8].copy())
        W = 4 * (res[0] + 2048 * res[1]).astype(np.int64)
        print(f"A K-major layout_type={lt} LBO=16 SBO=1024:")
        for r in (0, 1, 2, 7, 8, 9, 127):
            print(f"  bytes (m={r}, k=0..7):", W[r, :8].tolist())


GROUPS = dict(ts=group_ts, mn_swizzle=group_mn_swizzle, kmajor=group_kmajor, mn_discover=group_mn_discover, mn_check=group_mn_check, rounding=group_rounding, rate=group_rate)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        GROUPS[sys.argv[1]]()
    else:
        for g in GROUPS:
            print(f"==== {g}", flush=True)
            r = subprocess.run([sys.executable, __file__, g], capture_output=True, text=True, timeout=120)
            print(r.stdout + ("\n[stderr] " + r.stderr[-600:] if r.returncode or "umma_probe:" in r.stderr else ""), flush=True)

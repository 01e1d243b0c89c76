"""tcgen05 kind::f16 (bf16) probe driver (needs a B200): pins down the swizzled shared-memory layouts the
round-2 kernel family uses -- ONE tile X[line u][n] (bf16, 128 B or 32 B swizzle) read through a K-major
descriptor (lines = M/N index, n = reduction index: the weight-gradient GEMM) and through an MN-major
descriptor (n = N index, lines = reduction index: forward / data-gradient GEMM) -- and the issue rate per N.
Each group runs in its own process (a bad descriptor kills the CUDA context)."""
import ctypes as C
import subprocess
import sys

import numpy as np

sys.path.insert(0, ".")

LT = {128: 2, 64: 4, 32: 6}


def _lib():
    from pinn_based_online_pde_calculator_b200.engine import load_library

    lib = load_library()
    lib.pinn_umma_probe.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double),
                                    C.c_void_p]
    return lib


def bf16_bits(x):
    x = np.ascontiguousarray(x, np.float32)
    b = x.view(np.uint32)
    assert np.all((b & 0xFFFF) == 0), "values must be bf16-exact"
    return (b >> 16).astype(np.uint16)


def tile_image(X, sw):
    """X[U][N] -> swizzled image (uint16): n-block b (sw/2 elements) at b*U*sw, line u at u*sw, 16-byte chunk
    index XORed with address bits [7, 7+log2(sw/16))  (Swizzle<B,4,3> on the byte address)."""
    U, N = X.shape
    nb = sw // 2
    nblk = (N + nb - 1) // nb
    img = np.zeros(nblk * U * nb, np.uint16)
    u, n = np.meshgrid(np.arange(U), np.arange(N), indexing="ij")
    addr = (n // nb) * (U * sw) + u * sw + ((n % nb) // 8) * 16 + (n % 8) * 2
    mask = sw // 16 - 1
    addr = addr ^ (((addr >> 7) & mask) << 4)
    img[(addr // 2).ravel()] = bf16_bits(X).ravel()
    return img


def pad_words(img16, mult=256):
    w = img16.view(np.uint32) if img16.size % 2 == 0 else np.concatenate([img16, np.zeros(1, np.uint16)]).view(np.uint32)
    pad = (-w.size) % mult
    return np.concatenate([w, np.zeros(pad, np.uint32)])


def desc_k_major(U, sw):
    """lines = M/N index, n = K: (lbo, sbo, step, step2)"""
    if sw == 32:
        return (16, 8 * sw, U * sw, 0)
    return (16, 8 * sw, 32, U * sw)


def desc_mn_major(U, sw):
    """n = M/N index (contiguous), lines = K: (lbo, sbo, step, step2)"""
    return (U * sw, 8 * sw, 16 * sw, 0)


def run(lib, Aimg, Bimg, M, N, ksteps, a_mn, b_mn, ad, bd, a_lt, b_lt, reps=1, nd=1):
    A = pad_words(Aimg)
    B = pad_words(Bimg)
    cfg = np.array([M, N, ksteps, a_mn, b_mn, 1, reps, A.size, B.size, ad[0], ad[1], ad[2], bd[0], bd[1], bd[2], nd, a_lt, b_lt, 0, 1,
                    ad[3], bd[3]], np.int32)
    out = np.empty((128, 512), np.float32)
    cyc = C.c_double()
    st = np.zeros(2, np.int32)
    rc = lib.pinn_umma_probe(0, A.ctypes.data, B.ctypes.data, None, cfg.ctypes.data, out.ctypes.data, C.byref(cyc), st.ctypes.data)
    return rc, out, cyc.value, st


def ints(rng, *s):
    return rng.randint(-4, 5, size=s).astype(np.float32)


def group_fwd():
    """D[m][n] = sum_k W[m][k] * Y[k][n]: A = weights K-major image [M][K], B = activation tile [K lines][N], MN-major"""
    lib = _lib()
    rng = np.random.RandomState(0)
    for sw_a, sw_b in [(128, 128), (32, 32), (128, 32), (32, 128)]:
        for (N, K) in [(128, 64), (128, 128), (160, 128), (80, 128), (256, 64), (64, 32), (48, 128)]:
            if sw_b == 128 and N % 64 and N > 64:
                pass  # partial last block is fine: the image is padded
            Wm, Y = ints(rng, 128, K), ints(rng, K, N)
            rc, out, cyc, st = run(lib, tile_image(Wm, sw_a), tile_image(Y, sw_b), 128, N, K // 16, 0, 1, desc_k_major(128, sw_a),
                                   desc_mn_major(K, sw_b), LT[sw_a], LT[sw_b])
            ok = np.array_equal(out[:, :N], Wm @ Y)
            print(f"fwd  A K-major SW{sw_a}  B MN-major SW{sw_b}  N={N} K={K}: rc={rc} done={st[0]} match={ok}", flush=True)
            if not ok and st[0]:
                bad = np.argwhere(out[:, :N] != Wm @ Y)
                print("   first mismatches (m, n):", bad[:6].tolist(), " rows wrong:", len(set(bad[:, 0])), " cols wrong:", len(set(bad[:, 1])))


def group_wgrad():
    """D[i][j] = sum_n Y[i][n] * G[j][n]: both operands are activation tiles [unit lines][n] read K-major"""
    lib = _lib()
    rng = np.random.RandomState(1)
    for sw in (128, 32):
        for (Nj, R) in [(128, 128), (256, 128), (128, 160), (128, 80), (256, 64), (128, 48)]:
            Y, G = ints(rng, 128, R), ints(rng, Nj, R)
            ks = R // 16
            rc, out, cyc, st = run(lib, tile_image(Y, sw), tile_image(G, sw), 128, Nj, ks, 0, 0, desc_k_major(128, sw), desc_k_major(Nj, sw),
                                   LT[sw], LT[sw])
            ok = np.array_equal(out[:, :Nj], Y @ G.T)
            print(f"wgrad  both K-major SW{sw}  N={Nj} rows(K)={R}: rc={rc} done={st[0]} match={ok}", flush=True)
            if not ok and st[0]:
                bad = np.argwhere(out[:, :Nj] != Y @ G.T)
                print("   first mismatches (i, j):", bad[:6].tolist())


def group_rate():
    lib = _lib()
    rng = np.random.RandomState(2)
    K = 128
    for sw_a, sw_b in [(128, 128), (32, 32), (128, 32)]:
        for N in (64, 80, 96, 128, 160, 192, 256):
            Wm, Y = ints(rng, 128, K), ints(rng, K, N)
            t = {}
            nd = 2 if 2 * N <= 512 else 1
            for reps in (8, 40):
                rc, out, cyc, st = run(lib, tile_image(Wm, sw_a), tile_image(Y, sw_b), 128, N, K // 16, 0, 1, desc_k_major(128, sw_a),
                                       desc_mn_major(K, sw_b), LT[sw_a], LT[sw_b], reps=reps, nd=nd)
                t[reps] = cyc
            n = 32 * (K // 16)
            c = (t[40] - t[8]) / n
            print(f"rate fwd-shape A SW{sw_a} B SW{sw_b} N={N}: {c:.1f} cycles per tcgen05.mma (K=16) -> {2 * 128 * N * 16 / c:.0f} flop/cycle/SM", flush=True)
    # wgrad shape (both K-major)
    for sw in (128, 32):
        for Nj in (128, 256):
            R = 128
            Y, G = ints(rng, 128, R), ints(rng, Nj, R)
            t = {}
            for reps in (8, 40):
                rc, out, cyc, st = run(lib, tile_image(Y, sw), tile_image(G, sw), 128, Nj, R // 16, 0, 0, desc_k_major(128, sw),
                                       desc_k_major(Nj, sw), LT[sw], LT[sw], reps=reps, nd=1)
                t[reps] = cyc
            c = (t[40] - t[8]) / (32 * (R // 16))
            print(f"rate wgrad-shape SW{sw} N={Nj}: {c:.1f} cycles per tcgen05.mma (K=16)", flush=True)


def group_rounding():
    """how kind::f16 accumulates: 1 + 7 adds of 0.75 ulp across MMAs (RN +7ish, truncation 0)"""
    lib = _lib()
    K = 128
    for sign in (1.0, -1.0):
        Wm = np.zeros((128, K), np.float32)
        Y = np.zeros((K, 64), np.float32)
        Wm[:, 0] = sign
        Y[0, :] = 1.0
        for j in range(1, K // 16):
            Wm[:, 16 * j] = sign * 2.0 ** -12
            Y[16 * j, :] = 1.5 * 2.0 ** -12
        rc, out, cyc, st = run(lib, tile_image(Wm, 128), tile_image(Y, 128), 128, 64, K // 16, 0, 1, desc_k_major(128, 128),
                               desc_mn_major(K, 128), 2, 2)
        print(f"accumulate across MMAs, sign {sign:+.0f}: 7 adds of 0.75 ulp -> (D - sign)/ulp = {(out[0, 0] - sign) / 2.0 ** -23:+.2f}"
              f" (RN +-7, truncate 0, exact +-5.25; done={st[0]})")
    Wm = np.zeros((128, 16), np.float32)
    Y = np.zeros((16, 64), np.float32)
    Wm[:, 0] = 1.0
    Y[0, :] = 1.0
    Wm[:, 1:] = 2.0 ** -12
    Y[1:, :] = 1.5 * 2.0 ** -12
    rc, out, cyc, st = run(lib, tile_image(Wm, 128), tile_image(Y, 128), 128, 64, 1, 0, 1, desc_k_major(128, 128), desc_mn_major(16, 128), 2, 2)
    print(f"inside one MMA: 1 + 15 x 0.75 ulp -> (D-1)/ulp = {(out[0, 0] - 1) / 2.0 ** -23:+.2f} (exact 11.25)")
    rng = np.random.RandomState(3)
    for K in (128, 256):
        tb = lambda x: (x.view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)
        Wm = tb(rng.standard_normal((128, K)).astype(np.float32))
        Y = tb(rng.standard_normal((K, 64)).astype(np.float32))
        ks = K // 16
        ad = desc_k_major(128, 128)
        rc, out, cyc, st = run(lib, tile_image(Wm, 128), tile_image(Y, 128), 128, 64, ks, 0, 1, ad, desc_mn_major(K, 128), 2, 2)
        ref = Wm.astype(np.float64) @ Y.astype(np.float64)
        err = out[:, :64] - ref
        print(f"random bf16-exact inputs K={K}: rms rel = {np.sqrt((err ** 2).mean()) / np.sqrt((ref ** 2).mean()):.3e}, signed bias vs sign(ref) = "
              f"{(err * np.sign(ref)).mean() / np.abs(ref).mean():+.3e}  (done={st[0]})")


GROUPS = dict(fwd=group_fwd, wgrad=group_wgrad, rate=group_rate, rounding=group_rounding)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        GROUPS[sys.argv[1]]()
    else:
        for g in GROUPS:
            print(f"==== {g}", flush=True)
            try:
                r = subprocess.run([sys.executable, __file__, g], capture_output=True, text=True, timeout=180)
                print(r.stdout + ("\n[stderr] " + r.stderr[-800:] if r.returncode or "umma_probe:" in r.stderr else ""), flush=True)
            except subprocess.TimeoutExpired as e:
                print("TIMEOUT", (e.stdout or b"")[-2000:], flush=True)

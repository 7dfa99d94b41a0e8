"""Stand-alone probe (NOT product code): FP64 GEMM emulated on the INT8 tensor cores by Ozaki splitting, next to the
DMMA roofline this repository's kernels live under.

    C = A B,  A (m x k), B (k x n) in FP64.
    Row i of A is scaled by 2^-e_i (e_i = exponent of max_j |a_ij|), column j of B by 2^-f_j, so all entries lie in
    (-1, 1); each scaled entry is cut into S signed 7-bit digits  a = sum_s A_s 2^(-7 (s + 1)),  A_s in [-127, 127]
    (int8).  Every digit product A_s B_t is an EXACT int8 x int8 -> int32 GEMM (k * 127^2 < 2^31 for k <= 133 000), and
        C_ij = 2^(e_i + f_j) sum_{s + t < S} 2^(-7 (s + t + 2)) (A_s B_t)_ij
    keeps the S (S + 1) / 2 products whose weight is above the truncation level 2^(-7 S) (Ootomo, Ozaki, Yokota:
    "DGEMM on integer matrix multiplication unit", 2024).  The error is relative to max_i|a_i.| max_j|b_.j| per entry
    -- row / column scaling, NOT the entrywise |A||B| bound of FP64 -- so cancellation-heavy products (the Cholesky
    trailing update of an ill-conditioned kernel matrix is one) lose digits a true FP64 GEMM keeps; that error analysis,
    not the speed, is what stands between this probe and a product kernel.

The int8 GEMMs go through torch._int_mm (cuBLASLt): this measures what the tensor cores deliver for the scheme, not a
hand-written tcgen05 kernel.  Prints one JSON line per configuration.
"""
import json
import sys

import torch

dev = torch.device("cuda", 0)


def split(x, S, dim):
    """x (FP64) -> S int8 digit matrices and the per-row (dim=1) / per-column (dim=0) exponents."""
    amax = x.abs().amax(dim=dim, keepdim=True).clamp_min(1e-300)
    e = torch.ceil(torch.log2(amax)) + 1.0          # |x| 2^-e < 1/2: the first digit stays within +-63, carries fit
    r = x * torch.exp2(-e)
    digits = []
    for _ in range(S):
        r = r * 128.0
        d = torch.round(r)
        digits.append(d.to(torch.int8))
        r = r - d                                    # |r| <= 1/2 -> next digit within +-64
    return digits, e


def ozaki_gemm(A, B, S):
    Ad, ea = split(A, S, dim=1)
    Bd, fb = split(B, S, dim=0)
    C = torch.zeros((A.shape[0], B.shape[1]), dtype=torch.float64, device=A.device)
    n_gemm = 0
    for lvl in range(S - 1, -1, -1):                 # smallest weights first
        acc = None
        for s in range(lvl + 1):
            p = torch._int_mm(Ad[s], Bd[lvl - s])    # exact int32
            acc = p if acc is None else acc + p      # |sum| <= (lvl + 1) k 127^2 < 2^31 for k <= 16384, S <= 8
            n_gemm += 1
        C += acc.to(torch.float64) * 2.0 ** (-7 * (lvl + 2))
    return C * torch.exp2(ea) * torch.exp2(fb), n_gemm


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    g = torch.Generator(device=dev).manual_seed(0)
    cases = {}
    # (1) well-scaled random matrices
    cases["randn"] = (torch.randn((n, n), dtype=torch.float64, device=dev, generator=g),
                      torch.randn((n, n), dtype=torch.float64, device=dev, generator=g))
    # (2) rows spanning 12 orders of magnitude (what the row / column exponents are for)
    sc = torch.exp2(torch.randint(-20, 20, (n, 1), device=dev, generator=g).to(torch.float64))
    cases["row_scaled"] = (cases["randn"][0] * sc, cases["randn"][1] * sc.t())
    # (3) the trailing update of this repository's path: L L^T of an RBF kernel matrix factor (cond ~ 1e6), where
    #     K - L L^T cancels to rounding level
    t = torch.sort(torch.rand(n, dtype=torch.float64, device=dev, generator=g)).values
    K = 1.5 * torch.exp(-0.5 * ((t[:, None] - t[None, :]) / 0.05) ** 2) + 1e-2 * torch.eye(n, dtype=torch.float64, device=dev)
    L = torch.linalg.cholesky(K)
    cases["chol_factor_LLt"] = (L, L.t().contiguous())
    t_dgemm, _ = timed(lambda: cases["randn"][0] @ cases["randn"][1])
    for name, (A, B) in cases.items():
        ref = A @ B
        bound = (A.abs() @ B.abs()).clamp_min(1e-300)            # entrywise FP64 error scale
        rowcol = A.abs().amax(1, keepdim=True) * B.abs().amax(0, keepdim=True) * A.shape[1]
        for S in (5, 6, 7, 8):
            ms, (C, n_gemm) = timed(lambda: ozaki_gemm(A, B, S), reps=2)
            ms_mm, _ = timed(lambda: torch._int_mm(torch.ones((n, n), dtype=torch.int8, device=dev),
                                                   torch.ones((n, n), dtype=torch.int8, device=dev)), reps=3) if S == 5 and name == "randn" else (None, None)
            err = (C - ref).abs()
            out = {"case": name, "n": n, "slices": S, "int8_gemms": n_gemm, "ms": ms,
                   "fp64_equiv_tflops": 2.0 * n ** 3 / (ms * 1e-3) / 1e12,
                   "dgemm_cublas_ms": t_dgemm, "dgemm_cublas_tflops": 2.0 * n ** 3 / (t_dgemm * 1e-3) / 1e12,
                   "max_err_over_abs_product": float((err / bound).max()),
                   "max_err_over_rowmax_colmax_k": float((err / rowcol).max()),
                   "max_err_over_max_ref": float(err.max() / ref.abs().max())}
            if ms_mm is not None:
                out["one_int8_gemm_ms"] = ms_mm
                out["int8_tops"] = 2.0 * n ** 3 / (ms_mm * 1e-3) / 1e12
            if name == "chol_factor_LLt":
                out["max_err_of_K_minus_LLt_over_diagK"] = float(((K - C).abs().max()) / K.diagonal().max())
                out["fp64_K_minus_LLt_over_diagK"] = float(((K - ref).abs().max()) / K.diagonal().max())
            print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()

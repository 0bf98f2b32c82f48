"""Round-2 A/B measurements (one JSON line per run; run on a B200 under gpurun):

  python scripts/exp_r2.py spmv  <workload> [reps]    workload: c1 | c4 | c5s24 | c5shard (scale 27, 1/8 of the rows)
  python scripts/exp_r2.py libs  <workload> lib,lib,.. the spmv run once per build of the library
                                                       (SPBLAS_B200_LIB; "base" = the shipped one)

The library is chosen before import through SPBLAS_B200_LIB, so every build runs in its own
process; generated matrices are cached under /tmp for the later processes of the same call."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps, warm=3):
    import torch
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cached(name, make):
    import torch
    path = f"/tmp/exp_r2_{name}.pt"
    if os.path.exists(path):
        return torch.load(path, map_location="cuda:0")
    obj = make()
    try:
        torch.save(obj, path)
    except Exception:
        pass
    return obj


def spmv_workload(wl):
    import torch
    from spblas_reference_b200 import generators as G
    from spblas_reference_b200.sharded import balanced_nnz_blocks
    dev = torch.device("cuda:0")
    if wl == "c1":
        sets = []
        for c in range(4):
            v, rp, ci, shape = G.uniform_random_csr(1_000_000, 1_000_000, 10, seed=c,
                                                    dtype=torch.float32, device=dev)
            sets.append((v, rp, ci, shape, G.dense_uniform((shape[1],), 5, torch.float32, dev)))
        return sets
    if wl == "c4":
        v, rp, ci, shape = cached("c4", lambda: G.rmat_csr(24, 16, seed=24, dtype=torch.float32, device=dev))
        return [(v, rp, ci, tuple(shape), G.dense_uniform((shape[1],), 5, torch.float32, dev))]
    if wl == "c5s24":
        v, rp, ci, shape = cached("c5s24", lambda: G.rmat_csr(24, 16, seed=27, dtype=torch.float64,
                                                              device=dev, off_dtype=torch.int64))
        return [(v, rp, ci, tuple(shape), G.dense_uniform((shape[1],), 5, torch.float64, dev))]
    if wl == "c5shard":
        def make():
            scale, parts = 27, 8
            n = 1 << scale
            deg = G.rmat_degrees(scale, 16, seed=27, device=dev)
            rowptr_all = torch.zeros(n + 1, dtype=torch.int64, device=dev)
            torch.cumsum(deg, 0, out=rowptr_all[1:])
            blocks = balanced_nnz_blocks(rowptr_all, parts)
            del deg, rowptr_all
            r0, r1 = blocks[parts // 2]          # a middle block: rank 0's holds the densest rows
            return G.rmat_csr(scale, 16, seed=27, dtype=torch.float64, device=dev,
                              off_dtype=torch.int64, row_begin=r0, row_end=r1)
        v, rp, ci, shape = cached("c5shard", make)
        return [(v, rp, ci, tuple(shape), G.dense_uniform((shape[1],), 5, torch.float64, dev))]
    raise SystemExit(f"unknown workload {wl}")


def set_persisting_l2(mb):
    """cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize): the set-aside L2::evict_last lines live in"""
    import ctypes
    import torch
    torch.cuda.init()
    rt = ctypes.CDLL("libcudart.so.12")
    size = ctypes.c_size_t(0)
    rc = rt.cudaDeviceSetLimit(6, ctypes.c_size_t(mb << 20))
    rt.cudaDeviceGetLimit(ctypes.byref(size), 6)
    prop_max = torch.cuda.get_device_properties(0)
    print(json.dumps({"persisting_l2_request_mb": mb, "rc": rc, "granted_mb": size.value >> 20,
                      "L2_cache_size_mb": prop_max.L2_cache_size >> 20}), flush=True)


def run_spmv(wl, reps):
    import torch
    import spblas_reference_b200 as sb
    from spblas_reference_b200 import _cabi
    if os.environ.get("EXP_PERSIST_L2_MB"):
        set_persisting_l2(int(os.environ["EXP_PERSIST_L2_MB"]))
    sets = spmv_workload(wl)
    ops = []
    opt = os.environ.get("EXP_MATRIX_OPT", "0") == "1"      # wrap in matrix_opt (hub tables)
    force = os.environ.get("EXP_VARIANT")                   # force a kernel variant
    hub_cols = os.environ.get("EXP_HUB_COLS")
    for v, rp, ci, shape, x in sets:
        a = sb.csr_view(v, rp, ci, shape, int(ci.numel()))
        if opt:
            a = sb.matrix_opt(a)
        y = torch.empty(shape[0], dtype=v.dtype, device=v.device)
        info = sb.multiply_inspect(a, x, y)
        if force is not None:
            if int(force) >= 3:
                info.set_hub(True, int(hub_cols) if hub_cols else 0, 0)
            info.force_spmv_variant(int(force))
        ops.append((a, x, y, info))

    def fn(i):
        a, x, y, info = ops[i % len(ops)]
        sb.multiply_execute(info, a, x, y)
    ms = timed(fn, reps)
    a, x, y, info = ops[0]
    nnz = int(sets[0][2].numel())
    sT = v.element_size()
    nbytes = nnz * (sT + 4) + (shape[0] + 1) * rp.element_size() + shape[1] * sT + shape[0] * sT
    print(json.dumps({"exp": "spmv", "workload": wl, "lib": os.path.basename(_cabi.LIB_PATH),
                      "variant": info.spmv_variant, "matrix_opt": opt, "forced": force,
                      "hub_count": info.hub_count, "hub_ref_share": round(info.hub_refs / max(nnz, 1), 4),
                      "ms": round(ms, 4), "nnz": nnz,
                      "rows": shape[0], "cols": shape[1],
                      "gflops": round(2.0 * nnz / ms / 1e6, 1),
                      "alg_gbs": round(nbytes / ms / 1e6, 1),
                      "checksum": float(y.double().sum().item())}), flush=True)


def run_spmm_rmat(scale, k, dtype_name):
    """row kernel vs stream (ring) kernel on a power-law matrix with a narrow B (forced through
    SPBLAS_B200_SPMM_VARIANT): does the row-length histogram have to steer the choice?"""
    import torch
    import spblas_reference_b200 as sb
    from spblas_reference_b200 import generators as G
    dev = torch.device("cuda:0")
    dt = torch.float32 if dtype_name == "fp32" else torch.float64
    v, rp, ci, shape = G.rmat_csr(scale, 16, seed=24, dtype=dt, device=dev)
    a = sb.csr_view(v, rp, ci, shape, int(ci.numel()))
    B = G.dense_uniform((shape[1], k), 4, dt, dev)
    ref = None
    for forced in ("0", "1", None):
        if forced is None:
            os.environ.pop("SPBLAS_B200_SPMM_VARIANT", None)
        else:
            os.environ["SPBLAS_B200_SPMM_VARIANT"] = forced
        C = torch.empty((shape[0], k), dtype=dt, device=dev)
        info = sb.multiply_inspect(a, B, C)
        ms = timed(lambda i: sb.multiply_execute(info, a, B, C), 10)
        ref = C if ref is None else ref
        print(json.dumps({"exp": "spmm_rmat", "scale": scale, "k": k, "dtype": dtype_name, "forced": forced,
                          "variant": info.spmm_variant, "max_row_len": info.max_row_len,
                          "mean_row_len": round(a.nnz / shape[0], 2), "num_segments": info.num_segments,
                          "ms": round(ms, 4), "max_abs_diff_vs_first": (C - ref).abs().max().item()}), flush=True)
        info.close()


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "spmm_rmat":
        run_spmm_rmat(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
        sys.exit(0)
    if mode == "spmv":
        run_spmv(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
    elif mode == "libs":
        for lib in sys.argv[3].split(","):
            env = dict(os.environ)
            if lib != "base":
                env["SPBLAS_B200_LIB"] = os.path.join(ROOT, "spblas_reference_b200", f"libspblas_b200_{lib}.so")
            subprocess.run([sys.executable, os.path.abspath(__file__), "spmv", sys.argv[2]] + sys.argv[4:],
                           env=env, check=False)

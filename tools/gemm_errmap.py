"""Debug aid: where does dgemm_dev differ from torch?  python tools/gemm_errmap.py m n k beta"""
import sys, torch
sys.path.insert(0, ".")
import eigenexa_b200 as E
m, n, k, beta = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
dev = torch.device("cuda:0")
E.eigen_init(None, "C")
torch.manual_seed(1)
lda, ldb, ldc = m + 4, n + 2, m + 2
A = torch.zeros(k, lda, dtype=torch.float64, device=dev); A[:, :m] = torch.rand(k, m, dtype=torch.float64, device=dev) - 0.5
B = torch.zeros(k, ldb, dtype=torch.float64, device=dev); B[:, :n] = torch.rand(k, n, dtype=torch.float64, device=dev) - 0.5
C = torch.rand(n, ldc, dtype=torch.float64, device=dev); C0 = C.clone()
alpha = -1.0 if beta != 0 else 1.0
torch.cuda.synchronize()
E.dgemm_dev("N", "T", m, n, k, alpha, A.data_ptr(), lda, B.data_ptr(), ldb, beta, C.data_ptr(), ldc)
E.sync()
ref = alpha * (B[:, :n].T @ A[:, :m]) + beta * C0[:, :m]       # (n x m) = C^T
err = (C[:, :m] - ref).abs()
bad = (err > 1e-9).nonzero()
print("bad elements:", bad.shape[0], "of", m * n)
if bad.shape[0]:
    tn, tm = bad[:, 0] // 64, bad[:, 1] // 128
    tiles = sorted(set(zip(tm.tolist(), tn.tolist())))
    print("bad tiles (tm,tn):", len(tiles), tiles[:40])
    t0 = tiles[0]
    sel = (tm == t0[0]) & (tn == t0[1])
    rows = sorted(set((bad[sel][:, 1] % 128).tolist())); cols = sorted(set((bad[sel][:, 0] % 64).tolist()))
    print("first bad tile", t0, "count", int(sel.sum()), "rows", rows[:40], "cols", cols[:40])
    for q in range(3):
        i, j = int(bad[sel][q * 40, 0]), int(bad[sel][q * 40, 1])
        prod = float((B[:, i] * A[:, j]).sum())
        # partial products per 16-wide k chunk
        parts = [(float((B[c:c + 16, i] * A[c:c + 16, j]).sum())) for c in range(0, k, 16)]
        print("elem col", i, "row", j, "got", float(C[i, j]), "ref", float(ref[i, j]), "C0", float(C0[i, j]), "AB", prod)
        print("   got - alpha*AB - beta*C0 =", float(C[i, j]) - alpha * prod - beta * float(C0[i, j]))
        print("   chunk products", [round(x, 4) for x in parts])
    # per-tile count histogram
    import collections
    cnt = collections.Counter(zip(tm.tolist(), tn.tolist()))
    print("counts per bad tile:", sorted(cnt.values())[:50])
E.eigen_free()

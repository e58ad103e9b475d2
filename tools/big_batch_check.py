"""64-bit indexing at scale: one 4 Mi-row batch at the C2 shape (activation stash 2.7 G floats, past 2^31) against the same
rows processed as two halves.  python tools/big_batch_check.py [log2_rows]
Measured (round 1): forward and dz bit-identical; dw / db of whole vs halves 2.0e-5 / 2.7e-5 with at most 16 384 rows per
TMEM accumulation chain (6.1e-4 / 8.0e-4 with 1 Mi rows per chain; the default cap is now 4096 rows)."""
import sys

import torch

sys.path.insert(0, '.')
import bench  # noqa: E402
import nif_b200  # noqa: E402

dev = torch.device('cuda:0')
B = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 22)
net = nif_b200.NIFMultiScale(bench.CFG_S, bench.CFG_P, "float32", seed=0, device=dev)
eng = net.engine
g = torch.Generator(device=dev).manual_seed(0)
z = torch.rand(B, 32, generator=g, device=dev) - 0.5
x = torch.rand(B, 2, generator=g, device=dev) * 2 - 1
tgt = torch.rand(B, 1, generator=g, device=dev) * 2 - 1
w_h, b_h = net.w_h.detach(), net.b_h.detach()
packed = eng.pack(w_h, b_h)


def run(lo, hi, dw, db, beta):
    zz, xx, tt = z[lo:hi].contiguous(), x[lo:hi].contiguous(), tgt[lo:hi].contiguous()
    u, stash = eng.forward(zz, xx, packed, save=True)
    loss = torch.zeros(1, device=dev)
    dz = eng.mse_backward(zz, xx, packed, u, stash, tt, None, 1.0 / B, loss, dw, db, beta)
    return u, dz, loss


dw1, db1 = torch.empty_like(w_h), torch.empty_like(b_h)
u1, dz1, l1 = run(0, B, dw1, db1, 0.0)
dw2, db2 = torch.empty_like(w_h), torch.empty_like(b_h)
ua, dza, la = run(0, B // 2, dw2, db2, 0.0)
ub, dzb, lb = run(B // 2, B, dw2, db2, 1.0)
torch.cuda.synchronize()
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
out = {"rows": B, "stash_floats": eng.save_floats_per_row * B,
       "u_equal": bool(torch.equal(u1[: B // 2], ua) and torch.equal(u1[B // 2:], ub)),
       "dz_equal": bool(torch.equal(dz1[: B // 2], dza) and torch.equal(dz1[B // 2:], dzb)),
       "dw_rel": rel(dw2, dw1), "db_rel": rel(db2, db1), "loss_rel": abs(float(la + lb) - float(l1)) / abs(float(l1)),
       "finite": bool(torch.isfinite(dw1).all() and torch.isfinite(u1).all())}
print(out)
assert out["u_equal"] and out["dz_equal"] and out["finite"] and out["dw_rel"] < 5e-5 and out["db_rel"] < 5e-5 and out["loss_rel"] < 1e-5, out
print("big batch ok")

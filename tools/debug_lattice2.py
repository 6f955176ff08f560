import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cuda-fem_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import femx
from oracle import oracle as orc
ctx = femx.Context(0)
nx, ny, nz = 9, 7, 8
X, Y, Z, conn = orc.box_mesh(nx, ny, nz)
rng = np.random.RandomState(3)
X = X + rng.uniform(-0.01, 0.01, X.shape)
mesh = femx.Mesh(3, torch.from_numpy(conn).cuda(), tuple(torch.from_numpy(c).cuda() for c in (X, Y, Z)))
pat = femx.Pattern(ctx, mesh)
rp, ci = orc.pattern(conn, len(X))
idx = np.arange(len(X))
i, j, k = idx % (nx + 1), (idx // (nx + 1)) % (ny + 1), idx // ((nx + 1) * (ny + 1))
inner = (i > 0) & (i < nx) & (j > 0) & (j < ny) & (k > 0) & (k < nz)
rowof = np.repeat(idx, np.diff(rp))
for name, oid in (("POISSON", orc.POISSON), ("POISSON_MASS", orc.POISSON_MASS), ("MASS", orc.MASS)):
    ov = orc.assemble_csr(oid, 3, 1, conn, X, Y, Z, rp, ci, params=(1.0,))
    form = femx.Form(ctx, 3, getattr(femx, name), params=(1.0,))
    for lat in (1, 0):
        ctx.set_option("lattice", lat)
        v = form.assemble_csr(pat, mesh).cpu().numpy()
        d = np.abs(v - ov)
        m_in = inner[rowof]
        diag = ci == rowof
        print(name, "lattice", lat, "relF all %.3e" % (np.linalg.norm(v - ov) / np.linalg.norm(ov)),
              "interior rows %.3e" % (np.linalg.norm((v - ov)[m_in]) / np.linalg.norm(ov[m_in])),
              "boundary rows %.3e" % (np.linalg.norm((v - ov)[~m_in]) / np.linalg.norm(ov[~m_in])),
              "interior diag %.3e offdiag %.3e" % (np.linalg.norm((v - ov)[m_in & diag]) / np.linalg.norm(ov[m_in & diag]),
                                                   np.linalg.norm((v - ov)[m_in & ~diag]) / np.linalg.norm(ov[m_in & ~diag])),
              "worst", int(np.argmax(d)), "row", int(rowof[np.argmax(d)]), "col", int(ci[np.argmax(d)]), float(v[np.argmax(d)]), float(ov[np.argmax(d)]))
    form.close()

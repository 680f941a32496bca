#!/usr/bin/env python
"""Deviation of the fused tangent kernel from the Python rule, per parameter (prints one JSON line per lens)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from caustics_b200 import extended_source as es
g = np.load(os.path.join(ROOT, "tests", "golden", "ext_golden.npz"))
cases = [(2, dict(s=0.9, q=0.2), g["b_w_0.01"][:24]),
         (3, dict(s=0.9, q=0.2, q3=0.1, r3=0.8, psi=1.0), g["t_w_0.01"][:12])]
for nl, hp, w_np in cases:
    def run(fn):
        w = torch.from_numpy(w_np).cuda().requires_grad_(True)
        rho = torch.tensor(1e-2, dtype=torch.float64, device="cuda", requires_grad=True)
        t = {k: torch.tensor(v, dtype=torch.float64, device="cuda", requires_grad=True) for k, v in hp.items()}
        m = fn(w, rho, t)
        outs = []
        for i in range(len(w_np)):
            gr = torch.autograd.grad(m[i], [w, rho] + list(t.values()), retain_graph=True, allow_unused=True)
            outs.append([gr[0][i].real.item(), gr[0][i].imag.item()] + [x.item() for x in gr[1:]])
        return m.detach().cpu().numpy(), np.array(outs)
    for comp in (False, True):
        k = run(lambda w, rho, t: es._mag_uniform_kernel_grad(w, rho, nl, 200, 2500, comp, t))
        s = run(lambda w, rho, t: es._mag_from_contours(es._get_contours(w, rho, nl, 200, 2500, comp, t), w.reshape(-1), rho, nl, t))
        rel = np.abs(k[1] - s[1]) / (np.abs(s[1]) + 1e-9 * np.abs(s[1]).max(axis=0))
        print(json.dumps({"nl": nl, "compensated": comp, "mag_dev": float(np.max(np.abs(k[0] - s[0]) / s[0])),
                          "grad_dev_per_source": [float(x) for x in rel.max(axis=1)],
                          "grad_dev_per_param": [float(x) for x in rel.max(axis=0)]}))

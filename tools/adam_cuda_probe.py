#!/usr/bin/env python
"""Which float32 arithmetic does torch.optim.Adam run on CUDA?  (tests/test_gpu_kernels.py::test_adam_vs_torch_cuda_adam:
K3a follows torch's CPU kernels bit for bit; the CUDA foreach kernels group addcmul / addcdiv differently.)  One step of
torch.optim.Adam(foreach=True/False) on random state against candidate groupings evaluated on the host with exact
IEEE float32 semantics (fma through float64); prints the fraction of bit-identical elements per candidate."""
import json
import sys

import numpy as np
import torch

f32 = np.float32


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def main():
    rng = np.random.default_rng(5)
    shape = (1 << 20,)
    lr, b1, b2, eps = 0.005, 0.9, 0.999, 1e-8
    out = {}
    for foreach in (True, False):
        for t in (1, 3):
            g = (rng.standard_normal(shape) * 10.0 ** rng.uniform(-9, -2, size=shape)).astype(f32)
            m = (rng.standard_normal(shape) * 1e-5).astype(f32)
            v = (rng.random(shape) * 1e-9).astype(f32)
            p = (rng.standard_normal(shape) * 1e-2).astype(f32)
            P = torch.nn.Parameter(torch.from_numpy(p.copy()).cuda())
            opt = torch.optim.Adam([P], lr=lr, foreach=foreach)
            P.grad = torch.zeros_like(P)
            opt.step()                                                  # creates the state (step = 1) with a zero gradient
            st = opt.state[P]
            with torch.no_grad():
                P.copy_(torch.from_numpy(p)); st["exp_avg"].copy_(torch.from_numpy(m)); st["exp_avg_sq"].copy_(torch.from_numpy(v))
                st["step"].fill_(t - 1)
            P.grad = torch.from_numpy(g).cuda()
            opt.step()
            m_t, v_t, p_t = st["exp_avg"].cpu().numpy(), st["exp_avg_sq"].cpu().numpy(), P.detach().cpu().numpy()
            w1, B2, a2, ae = f32(1 - b1), f32(b2), f32(1 - b2), f32(eps)
            bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
            ss, bc2s = f32(-(lr / bc1)), f32(bc2 ** 0.5)
            eq = lambda a, b: float((a.view(np.int32) == b.view(np.int32)).mean())
            res = {}
            res["m: fma(w1, g-m, m)"] = eq(fma(np.full(shape, w1), g - m, m), m_t)
            vb = v * B2
            cand_v = {"fma(a2*g, g, v*b2)  [torch CPU]": fma(a2 * g, g, vb), "fma(a2, g*g, v*b2)": fma(np.full(shape, a2), g * g, vb),
                      "v*b2 + a2*(g*g)": vb + a2 * (g * g), "v*b2 + (a2*g)*g": vb + (a2 * g) * g,
                      "fma(g, a2*g, ...) == first": fma(g, a2 * g, vb)}
            for k, c in cand_v.items():
                res["v: " + k] = eq(c, v_t)
            sq = np.sqrt(v_t)
            dens = {"sqrt(v)/bc2s + eps": sq / bc2s + ae, "sqrt(v)*(1/bc2s in f32) + eps": sq * (f32(1) / bc2s) + ae,
                    "sqrt(v)*f32(1/double(bc2s)) + eps": sq * f32(1.0 / (bc2 ** 0.5)) + ae}
            for dk, den in dens.items():
                cand_p = {"p + (ss*m)/den  [torch CPU]": p + (ss * m_t) / den, "fma(ss, m/den, p)": fma(np.full(shape, ss), m_t / den, p),
                          "p + ss*(m/den)": p + ss * (m_t / den)}
                for k, c in cand_p.items():
                    res["p: den=%s ; %s" % (dk, k)] = eq(c, p_t)
            out["foreach=%s step=%d" % (foreach, t)] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

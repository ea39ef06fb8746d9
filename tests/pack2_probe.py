"""Helper of test_gpu_parity.py::test_packed_and_scalar_two_source_kernels_agree: runs the two-source golden
cases with the kernel form selected by MD2_PACK2 (read once per process by the library) and stores losses,
per-pixel gradients and masks in an .npz.  usage: pack2_probe.py out.npz"""
import sys

import numpy as np

from helpers import Golden
from gpu_driver import run_cuda

out = {}
for name in ("mono_iid", "mono_structured", "mono_jitterK", "disable_automasking", "no_ssim"):
    g = Golden(name)
    r = run_cuda(g, rows_per_segment=16)
    out[name + "/loss"] = np.array([float(r["losses"]["loss"].detach())] +
                                   [float(r["losses"]["loss/%d" % s].detach()) for s in range(4)])
    for s in range(4):
        out[name + "/gup%d" % s] = r["side"][("grad_updisp", s)].cpu().numpy()
        out[name + "/gd%d" % s] = r["leaves"][("disp", s)].grad.cpu().numpy()
        if g.n_id > 0:
            out[name + "/idsel%d" % s] = r["side"]["identity_selection/%d" % s].cpu().numpy()
    for f in g.frame_ids[1:]:
        out[name + "/gT%s" % f] = r["leaves"][("T", f)].grad.cpu().numpy()
        out[name + "/color%s" % f] = r["side"][("color", f, 0)].cpu().numpy()
    r0 = run_cuda(g, want_grad=False)
    out[name + "/loss_nograd"] = np.array([float(r0["losses"]["loss"])])
np.savez(sys.argv[1], **out)

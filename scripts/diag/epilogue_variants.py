"""Which (MRFA_CORR_CVT, MRFA_CORR_STORE) combinations of the correlation epilogue differ from (0, 0), where and by how much."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mrfa_b200 as m  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(33)
q = torch.randn(2, 256, 64, 64, device=dev) * 3.0
k = torch.randn(2, 256, 64, 64, device=dev) * 3.0
q[0, :, 0, 0] = 0.0
k[1, :, 5, 7] *= 1e-20
ref = None
for cvt in ("0", "1"):
    for store in ("0", "1", "2"):
        os.environ["MRFA_CORR_CVT"], os.environ["MRFA_CORR_STORE"] = cvt, store
        pyr = m.CorrPyramid(q, k, 256 ** -0.5)
        torch.cuda.synchronize()
        v = (pyr.volume0.clone(), pyr.volume1.clone())
        if ref is None:
            ref = v
            continue
        for lvl in (0, 1):
            a, b = v[lvl].view(torch.int16), ref[lvl].view(torch.int16)
            bad = (a != b).nonzero()
            msg = f"cvt={cvt} store={store} level{lvl}: {bad.shape[0]} of {a.numel()} differ"
            if bad.shape[0]:
                i = tuple(bad[0].tolist())
                msg += f"; first at {i}: got {v[lvl][i].item()!r} ({a[i].item() & 0xFFFF:#06x}) ref {ref[lvl][i].item()!r} ({b[i].item() & 0xFFFF:#06x})"
                rows = torch.unique(bad[:, 1])[:8].tolist()
                cols = torch.unique(bad[:, 2])[:8].tolist()
                msg += f"; rows {rows} cols {cols}; batches {torch.unique(bad[:, 0]).tolist()}"
            print(msg)

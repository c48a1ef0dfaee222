"""One launch of the persistent GEMV chain for ncu (32 x 4096^2 jobs, distinct activations)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mxq_b200 import ops  # noqa: E402
from profiles.r2_gemv_persistent import rand_packed, dev  # noqa: E402

oc, ic = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 4096)
share = int(sys.argv[3]) if len(sys.argv) > 3 else 0
n = 32
ps = [rand_packed(oc, ic) for _ in range(n)]
yy = [torch.empty(oc, device=dev, dtype=torch.float16) for _ in range(n)]
xs = [torch.randn(ic, device=dev).half() for _ in range(n)]
c = ops.GemvChain([((xs[0] if share else xs[i]), p, y, -1) for i, (p, y) in enumerate(zip(ps, yy))], validate=False)
for _ in range(3):
    c.run()
torch.cuda.synchronize()

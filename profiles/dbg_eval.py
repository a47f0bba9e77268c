import os, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from credgcn import evaluate
U, I, d = 31668, 38048, 64
dev = torch.device("cuda", 0)
torch.manual_seed(0)
fu = torch.randn(U, d, device=dev) * 0.1
fi = torch.randn(I, d, device=dev) * 0.1
csr = (torch.zeros(U + 1, dtype=torch.int64, device=dev), torch.zeros(1, dtype=torch.int32, device=dev))
users = torch.arange(U, device=dev)
for prec in ("bf16x3", "bf16"):
    for _ in range(2):
        evaluate.topk_device(fu, fi, users, csr, 20, prec)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); evaluate.topk_device(fu, fi, users, csr, 20, prec); b.record(); torch.cuda.synchronize()
    print(os.environ.get("CGX_OPT_EVAL_DEBUG", "0"), prec, round(a.elapsed_time(b), 3), "ms")

"""How long does one fused Adam launch chain take as a function of the parameter count?  (A sharded rank at N = 8 owns
~1.2 M parameters; its adam_kernel is still 30 us in the bench profile.)"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prodsearch_b200 import _lib
from prodsearch_b200.optimizers import Optimizer
for n in (4096, 65536, 1 << 20, 4 << 20, 16 << 20):
    p = torch.nn.Parameter(torch.randn(n, device="cuda"))
    opt = Optimizer("adam", 5e-4, 5.0)
    opt.set_parameters([("p", p)])
    p.grad = torch.randn(n, device="cuda") * 0.01
    for _ in range(5):
        opt.step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(50):
        opt.step()
    e.record()
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(10):
        opt.step()
    kp = _lib.profile_dump()
    _lib.profile_enable(False)
    print(json.dumps({"params": n, "us_per_step_3_launches": round(s.elapsed_time(e) / 50 * 1e3, 2),
                      "kernels_us": {k: round(v[1] / v[0] * 1e3, 2) for k, v in kp.items()}}), flush=True)

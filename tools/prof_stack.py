"""Developer tool: run the stack-mode kernel a few times for one size (for ncu)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from subpixal_b200 import batch
n = int(sys.argv[1]); npairs = int(sys.argv[2]); t = sys.argv[3] if len(sys.argv) > 3 else "NCC"
g = torch.Generator(device="cuda").manual_seed(1)
stack = torch.randn((npairs, 5, n, n), device="cuda", generator=g) + 3.0
for _ in range(4):
    out = batch.cc_fit_stack(stack, t)
torch.cuda.synchronize()
print(out[0][:2])

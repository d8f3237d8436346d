import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
buf = torch.zeros(2048, dtype=torch.int64, device="cuda")
os.environ["MM3D_TC_TRACE"] = str(buf.data_ptr())
sys.argv = ["run_conv_layer.py", "--kind", "smc", "--level", "0", "--cin", "16", "--cout", "16", "--mode", "tf32", "--dir", "fwd", "--reps", "1", "--noflush"]
import runpy
runpy.run_path(os.path.join(os.path.dirname(__file__), "run_conv_layer.py"), run_name="__main__")
torch.cuda.synchronize()
t = buf.cpu().view(-1, 4)
n = int((t[:, 3] != 0).sum())
t = t[:n]
print("items traced", n)
t0 = t[0, 0].item()
prev_end = t0
for i in range(min(n, 60)):
    a, b, c, d = [x.item() for x in t[i]]
    print(f"it {i:3d} start +{a - t0:7d} gap_from_prev {a - prev_end:6d} wait_b {b - a:6d} wait_a {c - b:6d} issue {d - c:6d}")
    prev_end = d
print("total cycles for traced items", t[n - 1, 3].item() - t0, "per item", (t[n - 1, 3].item() - t0) / n)

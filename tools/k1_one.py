"""Run selected K1 corruptions a few times (profiling target).  python tools/k1_one.py gaussian_noise impulse_noise ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fav
names = sys.argv[1:] or ["gaussian_noise"]
hw, n = 32, 65536
clf = fav.VisionClassifier("resnet18", 10, (hw, hw))
x = torch.randint(0, 256, (n, hw, hw, 3), dtype=torch.uint8, device="cuda")
out = torch.empty((n, hw, hw, 3), dtype=torch.bfloat16, device="cuda")
for name in names:
    cfg = fav.CorruptionConfig(None if name == "clean" else name, 0 if name == "clean" else 3)
    for _ in range(3):
        clf.corrupt_normalize(x, cfg, 0, 0, out=out)
torch.cuda.synchronize()
print("ok")

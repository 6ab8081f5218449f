"""ncu case: one configs[4] sweep point on the clusters-of-4 instantiation of the tensor-core solver
(B = 4096, H = 1024 by default: 64 tiles on 37 co-resident clusters)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from sweep import point
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
print(point(torch.device("cuda:0"), B, H, "dopri5", reps=1))

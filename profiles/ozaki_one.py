#!/usr/bin/env python
"""one int8-path product of the matvec shape (for ncu captures): python profiles/ozaki_one.py [slices=7] [reps=3]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tenpy_b200 import backend

s = int(sys.argv[1]) if len(sys.argv) > 1 else 7
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lib = backend.get_lib()
m, n, k = 2048, 4096, 1024
A = torch.randn(m * k, dtype=torch.float64, device=lib.device)
B = torch.randn(k * n, dtype=torch.float64, device=lib.device)
C = torch.empty(m * n, dtype=torch.float64, device=lib.device)
a_s = lib.ozaki_split(m, k, A, k, 1, s)
b_s = lib.ozaki_split(n, k, B, 1, n, s)
for _ in range(reps):
    lib.ozaki_mm(m, n, k, s, a_s, b_s, C, n)
torch.cuda.synchronize()
lib.ozaki_check_abort()
print('ok')

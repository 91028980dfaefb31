#!/usr/bin/env python
"""one block SVD (for ncu launch lists): python profiles/svd_one.py [n=1024] [variant=3]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tenpy_b200 import backend

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lib = backend.get_lib()
lib.svd_set_eig_variant(variant)
A = torch.randn(n * n, dtype=torch.float64, device=lib.device)
U, S, V = backend.zeros(n * n), backend.zeros(n), backend.zeros(n * n)
info, _, _ = lib.block_svd([n], [n], [0], [0], [0], [0], A, U, S, V)
torch.cuda.synchronize()
print('sweeps', info)

import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffsol_b200 import capi
L = capi.lib(); vp = ctypes.c_void_p; dev = torch.device("cuda:0")
n, B = 256, 2048
a = torch.randn((B, n, n), dtype=torch.float64, device=dev) + torch.eye(n, dtype=torch.float64, device=dev) * 4
rhs = torch.randn((B, n), dtype=torch.float64, device=dev)
piv = torch.zeros((B, n), dtype=torch.int32, device=dev); info = torch.zeros(B, dtype=torch.int32, device=dev)
for _ in range(2):
    capi.check(L.dsb_lu_factor_instance_major(vp(a.data_ptr()), n, B, vp(piv.data_ptr()), vp(info.data_ptr()), None))
    capi.check(L.dsb_lu_solve_instance_major(vp(a.data_ptr()), vp(piv.data_ptr()), vp(rhs.data_ptr()), n, B, vp(info.data_ptr()), None))
torch.cuda.synchronize()

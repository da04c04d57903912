import sys, ctypes, torch
ctypes.CDLL('scratch/libsegv_bt.so').install_segv_bt()
import custom_sparse_ops as cso
rowptr = torch.tensor([0,1,2], dtype=torch.int32, device='cuda')
col = torch.tensor([0,1], dtype=torch.int32, device='cuda')
vals = torch.ones(2, device='cuda')
try:
    cso.spmm_cpp.csr_spmm(rowptr, col, vals, 2, 2, torch.ones(3, 8, device='cuda'))
except RuntimeError as e:
    print('caught', str(e)[:100])
print('done')

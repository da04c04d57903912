import sys, torch, faulthandler
faulthandler.enable()
import custom_sparse_ops as cso
which = sys.argv[1]
rowptr = torch.tensor([0,1,2], dtype=torch.int32, device='cuda')
col = torch.tensor([0,1], dtype=torch.int32, device='cuda')
vals = torch.ones(2, device='cuda')
try:
    if which == 'direct_mismatch':
        cso.spmm_cpp.csr_spmm(rowptr, col, vals, 2, 2, torch.ones(3, 8, device='cuda'))
    elif which == 'direct_ok':
        print(cso.spmm_cpp.csr_spmm(rowptr, col, vals, 2, 2, torch.ones(2, 8, device='cuda')))
    elif which == 'fn_mismatch':
        a = torch.sparse_coo_tensor(torch.tensor([[0,1],[0,1]]), torch.ones(2), (2,2)).cuda().coalesce()
        cso.spmm(a, torch.ones(3, 8, device='cuda'))
    elif which == 'fn_ok':
        a = torch.sparse_coo_tensor(torch.tensor([[0,1],[0,1]]), torch.ones(2), (2,2)).cuda().coalesce()
        print(cso.spmm(a, torch.ones(2, 8, device='cuda')))
except RuntimeError as e:
    print('caught', str(e)[:100])
print(which, 'done')

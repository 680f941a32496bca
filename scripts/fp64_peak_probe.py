import torch, sys
sys.path.insert(0,'/root/repo')
from caustics_b200 import _lib
L=_lib.lib(); sink=torch.zeros(8,dtype=torch.float64,device='cuda'); st=torch.cuda.current_stream().cuda_stream
blocks,iters=148*8,3*(1<<14)
for name,fn in (('const-operand chains',L.caustics_bench_fp64_peak),('3 distinct operands',L.caustics_bench_fp64_peak3)):
    for _ in range(2): fn(sink.data_ptr(),blocks,iters,st)
    best=1e9
    for _ in range(5):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(sink.data_ptr(),blocks,iters,st); b.record(); torch.cuda.synchronize(); best=min(best,a.elapsed_time(b))
    print(name, 2.0*8*256*blocks*iters/(best*1e-3)/1e12, 'TFLOP/s')

import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from splat_b200 import _lib
W,H=1920,1080
sc=bench.make_scene(300000)
ctx=_lib.Context(device=0, lowpass=0.3); ctx.upload(sc)
cams=bench.orbit_cameras(W,H,2)
cs=_lib.camera_struct(bench._CamView(cams[0]))
dev=torch.device('cuda',0)
fb=torch.zeros((H,W),dtype=torch.int32,device=dev)
st=torch.cuda.current_stream()
for (r0,r1) in [(0,544),(544,1080)]:
    fb[r0:r1].zero_()
    ctx.render_device(cs, fb[r0:r1].data_ptr(), W,H,r0,r1, st.cuda_stream)
torch.cuda.synchronize()
host=torch.zeros((H,W),dtype=torch.int32).pin_memory()
host.copy_(fb, non_blocking=True); torch.cuda.synchronize()
hn=host.numpy().view(np.uint32)
print('device path checksum', int(hn.astype(np.uint64).sum()), 'nonzero', int(np.count_nonzero(hn)))
ref=np.zeros((H,W),np.uint32); ctx.render(cs, ref)
print('host path checksum', int(ref.astype(np.uint64).sum()), 'equal', bool(np.array_equal(ref,hn)))

import sys; sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch
import test_gpu_ops as t
from image_enhancement_deglaring_b200 import ops
dtype = ops.DG_F16
for cin, cout, H, W in [(8, 8, 16, 32), (8,8,32,32), (16,16,16,32)]:
    rs = t._rs(11); N = 1
    raw = torch.from_numpy((rs.standard_normal((N, cin, H, W)) * 3 + 1).astype(np.float32))
    q, seen = t._nhwc(raw, dtype)
    g, b = t._gn_params(rs, cin)
    w = torch.from_numpy((rs.standard_normal((cout, cin, 3, 3)) * (1.0 / np.sqrt(9 * cin))).astype(np.float32))
    wp = ops.pack_conv3x3(w.cuda()); wtc = ops.pack_conv3x3_tc(wp, dtype)
    st = t._stats(seen)
    src = ops.make_src(q, cin, stats=st, gamma=g.cuda(), beta=b.cuda(), groups=8)
    o_ref, s_ref = ops.conv3x3_fused([src], wp, cout, N, H, W, dtype, path=1)
    o_tc, s_tc = ops.conv3x3_fused([src], wp, cout, N, H, W, dtype, path=2, weight_tc=wtc)
    torch.cuda.synchronize()
    print(cin, cout, H, W, "out err", float((o_tc.float()-o_ref.float()).abs().max()))
    print(" ref", s_ref.flatten()[:8].cpu().numpy()); print(" tc ", s_tc.flatten()[:8].cpu().numpy())
    o = o_tc.float()[0]  # H W C
    print(" row sums ch0:", o[:, :, 0].sum(1)[:8].cpu().numpy(), " seg sums row0:", o[0,:16,0].sum().item(), o[0,16:,0].sum().item())

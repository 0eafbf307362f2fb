import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import hm_vae_b200 as H
from hm_vae_b200 import ops, _lib
from hm_vae_b200._lib import lib, ptr, stream, check
topo = json.load(open(os.path.join(ROOT, "tests", "golden", "topology.json")))["levels"]
lvl, ci, co, k, s, t_in, b = 3, 8, 16, 3, 1, 8, 4
nb = topo[lvl]["neighbours"]; j = len(nb)
torch.manual_seed(1)
conv = H.SkeletonConv(nb, j * ci, j * co, k, j, stride=s, padding=(k - 1) // 2, bias=True, padding_mode="reflection").cuda()
plan = conv.plan()
x = torch.randn(b, j * ci, t_in, device="cuda")
t_out = plan.t_out(t_in)
dy = torch.randn(b, j * co, t_out, device="cuda")
print("supported", lib.hmvae_conv_wgrad_tc_supported(plan.handle, b, t_in))
n = int(lib.hmvae_conv_wgrad_tc_workspace(plan.handle, b, t_in)); print("ws bytes", n)
ws = torch.zeros((n + 3) // 4, device="cuda")
gw = torch.zeros_like(conv.weight); gb = torch.zeros(j * co, device="cuda")
check(lib.hmvae_conv_wgrad_tc(plan.handle, ptr(x), ptr(dy), None, ptr(gw), ptr(gb), b, t_in, 0, ptr(ws), ws.numel() * 4, stream()), "wg")
torch.cuda.synchronize()
print("ws nonzero", int((ws != 0).sum()), "of", ws.numel())
print("gw nonzero", int((gw != 0).sum()), "of", gw.numel(), "abs sum", float(gw.abs().sum()))
ref = torch.zeros_like(conv.weight); rb = torch.zeros(j * co, device="cuda")
check(lib.hmvae_conv_wgrad(plan.handle, ptr(x), ptr(dy), None, ptr(ref), ptr(rb), b, t_in, 0, 1, stream()), "wg simt")
torch.cuda.synchronize()
print("ref nonzero", int((ref != 0).sum()), "abs sum", float(ref.abs().sum()), "rel", float((gw - ref).norm() / ref.norm()))
print("gw[0,0], ref[0,0]", gw[0, 0].tolist(), ref[0, 0].tolist())
print("gw[0,:4,0]", gw[0, :4, 0].tolist(), ref[0, :4, 0].tolist())

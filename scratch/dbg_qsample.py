import sys, torch
sys.path.insert(0, '.')
from tests import _golden as G
from oracle import skeldiff_oracle as oc
d = torch.device('cuda:0')
case = G.load_npz("amass_init")
spec, ae, diff, _, _ = G.dataset_models(case, device=d)
tabs = G.tables_of(case)
xq = diff.q_sample(case["x_start"].to(d), case["t_loss"].to(d), case["noise_loss"].to(d)).cpu()
ref = case["q_sample"]
print("t", case["t_loss"].tolist())
for b in range(xq.shape[0]):
    print(b, "err", float((xq[b]-ref[b]).abs().max()))
# alternatives
M = tabs["Umm_sqrt_Lambda_bar_t"]; t = case["t_loss"]
alt = tabs["sqrt_alphas_cumprod"][t][:,None,None]*case["x_start"] + M[t].transpose(1,2) @ case["noise_loss"]
print("transposed-M hypothesis err", float((xq-alt).abs().max()))
for tt in range(10):
    a = tabs["sqrt_alphas_cumprod"][tt]*case["x_start"][1] + M[tt] @ case["noise_loss"][1]
    print("sample1 as t=", tt, float((xq[1]-a).abs().max()))
print("buffer equal", torch.equal(diff.Umm_sqrt_Lambda_bar_t.cpu(), M), diff.Umm_sqrt_Lambda_bar_t.is_contiguous(), diff.Umm_sqrt_Lambda_bar_t.stride())

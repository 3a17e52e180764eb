import sys, torch
from collections import namedtuple
sys.path.insert(0, ".")
from cdgvae_b200.tabular.modules import model as M, train as T
from oracle import cdgvae_oracle as orc
kind = sys.argv[1] if len(sys.argv) > 1 else "adult"
n = 1 << 20
DS = namedtuple("DS", ["flatten_topology"]); Span = namedtuple("SpanInfo", ["dim", "activation_fn"])
torch.manual_seed(1)
if kind == "tvae":
    oil, mask, d, Bm, D = orc.tvae_shape("loan")
    cfg = dict(dataset="loan", scm="linear", flow_num=1, inverse_loop=100, node=d, factor=[1] * d, input_dim=D, sigma_range=[0.01, 0.1], lr=1e-3, weight_decay=1e-5)
    cfg["lambda"] = 5.0
    model = M.TVAE(Bm, mask, cfg, "cpu").to("cuda"); opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)
    x, y, nz = orc.synth_tvae("loan", n)
    run = lambda data: T.train_TVAE([[Span(*s) for s in col] for col in oil], None, data, model, cfg, opt, "cuda")
elif kind == "covtype":
    cfg = dict(dataset="covtype", scm="linear", flow_num=1, inverse_loop=100, lr=0.01, beta=0.01, node=6, factor=[1] * 6, input_dim=8)
    cfg["lambda"] = 10.0
    model = M.CDGVAE(orc.tabular_B("covtype"), [1, 1, 2, 1, 1, 8], cfg, "cpu").to("cuda"); opt = torch.optim.Adam(model.parameters(), lr=0.01)
    x, y, nz = orc.synth_tabular("covtype", n)
    run = lambda data: T.train_CDGVAE(DS(None), data, model, cfg, opt, "cuda")
else:
    cfg = dict(dataset="adult", scm="linear", flow_num=1, inverse_loop=100, lr=0.01, beta=0.01, node=3, factor=[1, 1, 1], input_dim=5)
    cfg["lambda"] = 10.0
    model = M.CDGVAE(orc.tabular_B("adult"), [1, 1, 3], cfg, "cpu").to("cuda"); opt = torch.optim.Adam(model.parameters(), lr=0.01)
    x, y, nz = orc.synth_tabular("adult", n)
    run = lambda data: T.train_CDGVAE(DS([2, 3, 0, 1, 4]), data, model, cfg, opt, "cuda")
nd = nz.cuda(); model.noise_fn = lambda a, b: nd
run([(x.cuda(), y.cuda())] * 3)
torch.cuda.synchronize(); print("ok")

import os, sys, torch
sys.path.insert(0, os.getcwd())
import bench
import multimodal_seq2seq_gscan_b200 as pkg
from multimodal_seq2seq_gscan_b200 import synthetic as O
cfg = bench.bench_cfg()
dev = torch.device("cuda:0")
model = pkg.Model(**O.model_kwargs(cfg)).to(dev)
model.load_state_dict(O.full_state_dict(O.synthetic_params(cfg, 1234)), strict=True)
host = bench.make_host_batch(cfg, 1235)
c = torch.tensor(host["commands"], device=dev); s = torch.tensor(host["situations"], device=dev); t = torch.tensor(host["targets"], device=dev)
model.train()
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    with torch.no_grad():
        logp, _ = model(commands_input=c, commands_lengths=host["cmd_lengths"], situations_input=s, target_batch=t, target_lengths=host["tgt_lengths"])
    torch.cuda.synchronize()

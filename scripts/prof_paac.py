"""Kernel-time breakdown of one PAAC update (torch.profiler), expanded vs compact observation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
import golds_rl_gym_b200 as pkg
import train_paac_conv as tp
from torch.profiler import profile, ProfilerActivity

E = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
for compact in (False, True):
    args = tp.get_arg_parser().parse_args(["--clip_norm=1", "-ec", str(E)])
    nc, ec = tp.get_network_and_environment_creator(args)
    L = pkg.submodule("agents.paac.paac").GridPAACLearner(nc, ec, args, use_cuda_graph=False, compact_obs=compact)
    L.start()
    for _ in range(3): L.update()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(2): L.update()
        torch.cuda.synchronize()
    print("==== compact =", compact)
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
    del L
    torch.cuda.empty_cache()

import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import mq_oracle as O, params as PR
from oracle.gen_golden import vilco_cfg
from util import rel_max
from vilco_b200.config import mq_model_kwargs
from vilco_b200.modeling import make_meta_arch

def run(prompts, adapters, tl):
    cfg = vilco_cfg()
    if not prompts: cfg.prompt_pool = None
    if not adapters: cfg.adapt_blocks, cfg.n_emas = (), 0
    P = PR.random_state(PR.param_spec(vilco_cfg()), 1)
    kw = mq_model_kwargs(cfg.input_dim, cfg.embd_dim, cfg.n_head, cfg.max_seq_len, cfg.arch, cfg.num_classes, cfg.n_txt_in, cfg.regression_range)
    if prompts: kw["cl_cfg"].update(name="l2p", prompt_pool=True, pool_size=10, topk=4, length=20, embed_dim=cfg.n_txt_in)
    if adapters: kw["cl_cfg"].update(use_adapt=True, adapt_blocks=[0, 1, 2, 3, 4])
    model = make_meta_arch("LocPointTransformer", **kw)
    model.load_state_dict(P, strict=False)
    model = model.cuda().eval()
    vids = PR.synth_video_list(cfg, 1, seed=5, lens=[900], text_lens=[tl], n_gt=[3])
    cls_l, _, _ = model(vids, is_training=False, get_emb=True)
    got = torch.cat(cls_l, 1)[0].cpu()
    with torch.no_grad():
        res, raw = O.model_infer(P, cfg, vids, return_raw=True)
    lg = torch.cat(raw[0][0], 1)[0]
    offs = [0]
    for l in raw[0][0]:
        offs.append(offs[-1] + l.shape[1])
    print(f"prompts={prompts} adapters={adapters} textlen={tl}", [round(rel_max(got[offs[i]:offs[i + 1]], lg[offs[i]:offs[i + 1]]), 5) for i in range(len(offs) - 1)], flush=True)

run(False, False, 57)
run(False, False, 120)
run(True, False, 57)
run(False, True, 57)

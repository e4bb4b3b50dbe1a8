import numpy as np
import torch

TOL = 1e-3  # north-star tolerance: logits / offsets / losses within 1e-3 relative (bf16 operands, fp32 accumulate)


def rel_max(a, b):
    """max |a - b| / max |b| — the relative error used throughout the parity tests."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def build_pair(cfg, seed=0):
    """(vilco_b200 model on cuda with seeded weights, the same weights as an oracle params dict)."""
    from oracle import params as PR
    from vilco_b200.config import mq_model_kwargs
    from vilco_b200.modeling import make_meta_arch
    P = PR.random_state(PR.param_spec(cfg), seed)
    model = make_meta_arch("LocPointTransformer", **mq_model_kwargs(
        cfg.input_dim, cfg.embd_dim, cfg.n_head, cfg.max_seq_len, cfg.arch, cfg.num_classes, cfg.n_txt_in,
        cfg.regression_range))
    missing, unexpected = model.load_state_dict(P, strict=False)
    assert not unexpected
    return model.cuda().eval(), P


def build_vilco_pair(cfg, seed=1):
    """mq_vilco.yaml-like model (L2P prompts, temporal adapters + EMA copy) with seeded weights."""
    from oracle import params as PR
    from vilco_b200.config import mq_model_kwargs
    from vilco_b200.modeling import make_meta_arch
    P = PR.random_state(PR.param_spec(cfg), seed)
    kw = mq_model_kwargs(cfg.input_dim, cfg.embd_dim, cfg.n_head, cfg.max_seq_len, cfg.arch, cfg.num_classes, cfg.n_txt_in,
                         cfg.regression_range)
    kw["cl_cfg"].update(name="l2p", memory_size=1010, prompt_pool=True, pool_size=cfg.prompt_pool["pool_size"],
                        topk=cfg.prompt_pool["top_k"], length=cfg.prompt_pool["length"], embed_dim=cfg.n_txt_in,
                        narration_ssl=True, use_adapt=True, adapt_blocks=list(cfg.adapt_blocks))
    model = make_meta_arch("LocPointTransformer", **kw)
    missing, unexpected = model.load_state_dict(P, strict=False)
    assert not unexpected
    return model.cuda().eval(), P


def build_vilco_train_pair(cfg, seed=2):
    """mq_vilco.yaml training model (prompts, adapters, narration SSL) with every dropout off — the configuration
    tests/golden/train_vilco.npz was generated with (oracle/gen_golden.py gen_vilco_train_golden)."""
    from oracle import params as PR
    from vilco_b200.config import mq_model_kwargs
    from vilco_b200.modeling import make_meta_arch
    P = PR.random_state(PR.param_spec(cfg), seed)
    kw = mq_model_kwargs(cfg.input_dim, cfg.embd_dim, cfg.n_head, cfg.max_seq_len, cfg.arch, cfg.num_classes, cfg.n_txt_in,
                         cfg.regression_range)
    kw["cl_cfg"].update(name="l2p", prompt_pool=True, pool_size=cfg.prompt_pool["pool_size"], topk=cfg.prompt_pool["top_k"],
                        length=cfg.prompt_pool["length"], embed_dim=cfg.n_txt_in, narration_ssl=True,
                        narration_dim=cfg.narration_dim, memory_size=48, ssl_factor=0.01, use_adapt=True,
                        adapt_blocks=list(cfg.adapt_blocks))
    kw["train_cfg"].update(dropout=0.0, droppath=1e-12)
    model = make_meta_arch("LocPointTransformer", **kw)
    missing, unexpected = model.load_state_dict(P, strict=False)
    assert not unexpected
    model.xl_dropout = 0.0
    return model.cuda(), P

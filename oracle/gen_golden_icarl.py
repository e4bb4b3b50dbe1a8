"""TEST INFRASTRUCTURE ONLY — golden vectors for the iCaRL nearest-mean-of-exemplars re-scoring (`PtTransformer.classify`,
MQ/libs/modeling/meta_archs.py:1061-1131, consumed by `inference` :1559-1562 and `inference_single_video` :1625-1643),
produced by running the REFERENCE ITSELF (oracle/ref_shim.py) with seeded weights, a seeded exemplar memory and a stand-in
for the dataset-side `cilsettask` (only its `get_dataloader(data_class, sample_frame=True)` is used).

Run in the authoring container only:   python -m oracle.gen_golden_icarl      -> tests/golden/icarl_small.npz
"""
import os

import numpy as np
import torch

from . import mq_oracle as O
from . import params as PR
from .gen_golden import GOLDEN, build_reference_model


def icarl_cfg():
    """`classify` hard-codes 10 pyramid levels (meta_archs.py:1065), so the golden model keeps the full-depth architecture
    [2, 2, 9] at T = 1024 and is small only in width."""
    return O.ModelCfg(input_dim=192, embd_dim=256, n_head=4, max_seq_len=1024, arch=(2, 2, 9), num_classes=6, n_txt_in=96)


class FakeTask:
    """what `classify` needs from QILSetTask: one loader per class, yielding video_list batches of one exemplar."""

    def get_dataloader(self, data_class, sample_frame=True):
        (videos,) = data_class.values()
        return [[v] for v in videos]


def memory_and_clips(c):
    """exemplar memory {class_id: [video dict, ...]} for EVERY class of the head (the re-scoring indexes the distance
    table with the flattened (T, num_classes) logit index, so both must have the same width) and two test clips."""
    T = c.max_seq_len
    ex = PR.synth_video_list(c, 2 * c.num_classes, seed=21, lens=[T] * (2 * c.num_classes),
                             text_lens=[20 + 3 * i for i in range(2 * c.num_classes)], n_gt=[2] * (2 * c.num_classes))
    memory = {k: [ex[2 * k], ex[2 * k + 1]] for k in range(c.num_classes)}
    clips = PR.synth_video_list(c, 2, seed=22, lens=[T, T - 28], text_lens=[31, 44], n_gt=[2, 3])
    return memory, clips


def main():
    c = icarl_cfg()
    model, _ = build_reference_model(c, seed=0)
    memory, clips = memory_and_clips(c)
    out = {}
    with torch.no_grad():
        for i, clip in enumerate(clips):
            model.memory = memory
            model.compute_means = True
            dists = model.classify(clip, FakeTask())                       # list over levels of (1, T_l, n_classes)
            for l, d in enumerate(dists):
                out[f"dists_{i}_{l}"] = d[0].numpy()
            if i == 0:      # the exemplar means themselves are 12 MB; keep their per-level, per-class norms and a few entries
                for l in range(len(dists)):
                    m = torch.stack(model.exemplar_means[l], 0)                               # (n_classes, C, T_l)
                    out[f"means_norm_{l}"] = m.flatten(1).norm(dim=1).numpy()
                    out[f"means_head_{l}"] = m[:, :4, :2].numpy()
            # the whole evaluation call: classify runs because compute_means is (again) True and a task object is passed
            model.compute_means = True
            res = model([clip], is_training=False, val_qilDatasetList=FakeTask())[0]
            assert model.compute_means is False
            out[f"det_segments_{i}"] = res["segments"].numpy()
            out[f"det_scores_{i}"] = res["scores"].numpy()
            out[f"det_labels_{i}"] = res["labels"].numpy()
            # inputs of the candidate selection, for the CPU test of the glue
            logits, offs, masks = model([clip], is_training=False, get_emb=True)
            for l in range(len(logits)):
                out[f"logits_{i}_{l}"] = logits[l][0].numpy()
                out[f"offsets_{i}_{l}"] = offs[l][0].numpy()
                out[f"mask_{i}_{l}"] = masks[l][0].numpy()
            plain = model([clip], is_training=False)[0]
            out[f"plain_scores_{i}"] = plain["scores"].numpy()
    np.savez_compressed(os.path.join(GOLDEN, "icarl_small.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.startswith(("dists_0", "det_", "means_head_0", "plain"))})
    print("size", os.path.getsize(os.path.join(GOLDEN, "icarl_small.npz")))


if __name__ == "__main__":
    main()

"""Default MQ configuration values needed to call `make_meta_arch` without the reference's yaml loader
(mirrors DEFAULTS / mq_no_cl.yaml of MQ/libs/core/config.py and MQ/configs/mq_no_cl.yaml)."""
import copy

TRAIN_CFG = dict(center_sample="radius", center_sample_radius=1.5, loss_weight=1.0, cls_prior_prob=0.01,
                 init_loss_norm=100, clip_grad_l2norm=1.0, head_empty_cls=[], dropout=0.1, droppath=0.1,
                 label_smoothing=0.0, t_c_alpha=0.8, use_dcn=False, dcn_start_layer=-1, use_us_fpn=False,
                 al_loss_weight=0.2, cont_loss_weight=0.0, seg_loss_weight=0.0, imp_loss_weight=0.0, temperature=0.07,
                 queue_size=256, length_theta=0.2, use_trident_head=False, num_bins=16, iou_weight_power=1.0)
TEST_CFG = dict(pre_nms_thresh=0.001, pre_nms_topk=5000, iou_threshold=0.1, min_score=0.0001, max_seg_num=200,
                nms_method="soft", nms_sigma=0.99, duration_thresh=0.01, multiclass_nms=True, ext_score_file=None,
                voting_thresh=0.9)
CL_CFG = dict(name=None, memory_size=0, random_order=False, reg_lambda=3000, type_sampling="icarl", adv_lambda=0,
              prompt_pool=False, pool_size=0, topk=4, length=20, embed_dim=768, narration_ssl=False, narration_dim=512,
              ssl_factor=0.01, use_adapt=False, adapt_blocks=[])


def mq_model_kwargs(input_dim=4096, embd_dim=1024, n_head=16, max_seq_len=1024, arch=(2, 2, 9), num_classes=22,
                    n_txt_in=768, regression_range=None, use_cross_modal=True, **over):
    """kwargs for make_meta_arch('LocPointTransformer', **kw) equal to load_config(mq_no_cl.yaml)['model']."""
    if regression_range is None:
        regression_range = [[0, 4], [2, 8], [4, 16], [8, 32], [16, 64], [32, 128], [64, 256], [128, 512], [256, 1024],
                            [512, 10000]]
    kw = dict(backbone_type="convTransformer", fpn_type="identity", use_xl=True, backbone_arch=list(arch), scale_factor=2,
              input_dim=[input_dim], max_seq_len=max_seq_len, max_buffer_len_factor=1.0, n_head=n_head, n_mha_win_size=-1,
              embd_kernel_size=3, embd_dim=[embd_dim], embd_with_ln=True, fpn_dim=embd_dim, fpn_with_ln=True,
              fpn_start_level=0, head_dim=embd_dim, regression_range=regression_range, head_num_layers=3,
              head_kernel_size=3, head_with_ln=True, use_abs_pe=True, use_rel_pe=False, num_classes=num_classes,
              train_cfg=copy.deepcopy(TRAIN_CFG), test_cfg=copy.deepcopy(TEST_CFG), cl_cfg=copy.deepcopy(CL_CFG),
              use_cross_modal=use_cross_modal, n_txt_in=n_txt_in)
    kw.update(over)
    return kw

"""lighthand_b200 -- B200-native (sm_100a) heatmap pipeline of LightHand.

Layout mirrors the reference's modules for the hot path only:

    lighthand_b200.utils.loss            JointsMSELoss, get_max_preds, PCK_2d_loss*, EPE, EPE_train
    lighthand_b200.utils.preprocessing   generate_patch_image, gen_trans_from_patch_cv, trans_point2d
    lighthand_b200.utils.metric_logger   AverageMeter, DeviceMeter
    lighthand_b200.utils.argparser       pred_store, pred_eval, pred_test, write_pck_eval (PCK curve / AUC / EPE sweep)
    lighthand_b200.utils.method          Runner_t, Runner_v (device-resident train / valid loops)
    lighthand_b200.utils.dir             dump (evaluation.json wire format)
    lighthand_b200.utils.dataset_loader  GenerateHeatmap
    lighthand_b200.tools.dataset         generate_target, GpuCollate
    lighthand_b200.ops                   batched device ops (one call = one kernel)
    lighthand_b200.pipeline              HeatmapPipeline (device-resident, CUDA graph), HostPipeline (C-ABI ctx)
    lighthand_b200.distributed           metric all-reduce across ranks

All compute runs in liblighthand_b200.so (hand-written CUDA); there is no CPU fallback.
"""
__version__ = "0.1.0"

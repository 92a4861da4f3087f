"""btcs_pnes_optical_flow_b200 -- B200-native hot path of saitosatoshi-1/BTCS_PNES_optical_flow.

Dense Farneback optical flow per frame pair -> body-axis projection -> ROI means (flow.py, mirrors the
reference's optical_flow.py), and sliding-window 2x2 PCA -> PC1 (pca.py, mirrors optical_PCA.py), both in
hand-written sm_100a CUDA behind a C-ABI (include/btcsflow.h, libbtcsflow.so).  PC1 metrics stay on the
host (metrics.py, mirrors optical_PC1.py).
"""
from .flow import (FB_PARAMS, FlowPlan, BtcsFlowError, Cv2CompatError, OPTFLOW_FARNEBACK_GAUSSIAN,
                   OPTFLOW_USE_INITIAL_FLOW, bgr_to_gray, build_roi_mask, calcOpticalFlowFarneback, clear_plans,
                   compute_roi_mean_body_flow, get_plan, run_body_axis_flow_core, skel_index_from_time)
from .pca import (bandpass_nanrobust, bandpass_nanrobust_device, butter_bandpass_sos, dynamic_pc1_sliding, flow_to_pc1,
                  pc1_sliding_batched)
from .metrics import compute_pc1_metrics

__version__ = "0.1.0"

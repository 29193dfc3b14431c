"""msr3d_b200 — B200-native hot path of MSR3D (PointNet++ object encoder + situated
spatial-attention transformer) behind the reference's module / extension API.

    from msr3d_b200 import OSE3DSituation, PcdObjEncoder      # drop-in nn.Modules
    import msr3d_b200.compat as compat; compat.install_ext()   # serve `pointnet2._ext`
    from msr3d_b200.splice import splice_scene_tokens           # llm_proj + scatter into the LLM input (msr3d.py:277-287)
    from msr3d_b200.data import preprocess_objects              # per-object preprocessing on the device (msr3d.py:181-215)
    from msr3d_b200.train import HotPathTrainer                 # fused, graph-replayed, pipelined training step

The compute runs in hand-written sm_100a kernels (msr3d_b200/csrc) reached through the C ABI
of include/msr3d_b200.h; importing this package never loads the oracle and never falls back to
CPU — a missing libmsr3d_b200.so raises on first use."""
from .compat import register_plugins as _register_plugins
from .encoder import PcdObjEncoder, PointNetPP, PointnetSAModule, PointnetSAModuleMSG, SharedMLP  # noqa: F401
from .prompter import OSE3DSituation  # noqa: F401
from .registry import (MODEL_REGISTRY, VISION_REGISTRY, build_model, build_module)  # noqa: F401
from .spatial import (MultiHeadAttentionSpatial, TransformerSpatialEncoderLayer,  # noqa: F401
                      calc_pairwise_locs, transform_to_agent_coor)

_register_plugins()
__version__ = "0.1.0"

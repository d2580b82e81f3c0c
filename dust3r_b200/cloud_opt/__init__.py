"""global_aligner(): drop-in for dust3r/cloud_opt/__init__.py:14-33."""
from enum import Enum

from .optimizer import PointCloudOptimizer
from .modular_optimizer import ModularPointCloudOptimizer
from .pair_viewer import PairViewer


class GlobalAlignerMode(Enum):
    PointCloudOptimizer = "PointCloudOptimizer"
    ModularPointCloudOptimizer = "ModularPointCloudOptimizer"
    PairViewer = "PairViewer"


_MODES = {
    GlobalAlignerMode.PointCloudOptimizer: PointCloudOptimizer,
    GlobalAlignerMode.ModularPointCloudOptimizer: ModularPointCloudOptimizer,
    GlobalAlignerMode.PairViewer: PairViewer,
}


def global_aligner(dust3r_output, device, mode=GlobalAlignerMode.PointCloudOptimizer, **optim_kw):
    view1, view2, pred1, pred2 = [dust3r_output[k] for k in 'view1 view2 pred1 pred2'.split()]
    if mode not in _MODES:
        raise NotImplementedError(f'Unknown mode {mode}')
    return _MODES[mode](view1, view2, pred1, pred2, **optim_kw).to(device)

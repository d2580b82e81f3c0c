"""Entry point of the alignment path: `global_aligner(dust3r_output, device, mode, **kw)` builds the optimizer object
for the output of inference() and moves it to `device` (dust3r/cloud_opt/__init__.py:14-33).  The classes keep the
reference's names and methods; `compute_global_alignment` runs the fused CUDA step (csrc/align_step.cu)."""
from enum import Enum

from .modular_optimizer import ModularPointCloudOptimizer
from .optimizer import PointCloudOptimizer
from .pair_viewer import PairViewer


class GlobalAlignerMode(Enum):
    PointCloudOptimizer = "PointCloudOptimizer"
    ModularPointCloudOptimizer = "ModularPointCloudOptimizer"
    PairViewer = "PairViewer"

    @property
    def optimizer_class(self):
        return {'PointCloudOptimizer': PointCloudOptimizer, 'ModularPointCloudOptimizer': ModularPointCloudOptimizer,
                'PairViewer': PairViewer}[self.value]


def global_aligner(dust3r_output, device, mode=GlobalAlignerMode.PointCloudOptimizer, **optim_kw):
    if not isinstance(mode, GlobalAlignerMode):
        raise NotImplementedError(f'Unknown mode {mode}')
    scene = mode.optimizer_class(dust3r_output['view1'], dust3r_output['view2'], dust3r_output['pred1'], dust3r_output['pred2'],
                                 **optim_kw)
    # non_blocking: predictions that sit in pinned host memory (what inference() returns) are uploaded asynchronously on the
    # current stream, under the rest of the host-side set-up; pageable sources copy synchronously as before
    return scene.to(device, non_blocking=True)

"""`import pyDeform` drop-in (reference: src/interface/pydeform.cc). Re-exports meshode_b200.pyDeform."""
import torch  # noqa: F401  (the reference requires torch to be imported first; README.md:46-50)

from meshode_b200.pyDeform import *  # noqa: F401,F403
from meshode_b200.pyDeform import (CeresEdges, CeresProblem, CeresSolve, DestroyTemplate,  # noqa: F401
                                   DistanceFieldLoss_forward_backward, EdgeLoss_backward_atomic, GetGrid, GetTemplateInfo,
                                   GridViews, LossForwardBackward, NearestVertex, SetGrid)

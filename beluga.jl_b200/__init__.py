"""beluga.jl_b200 — B200-native drop-in for the gene-family likelihood hot path of arzwa/Beluga.jl.

Python mirror of the reference's Julia API for that path (see INTEGRATION.md for the Julia `ccall`
side).  The model graph stays on the host; every number (ε, W, pruning recursion, root integration,
reductions) is computed by the CUDA library behind include/beluga_b200.h.  No CPU fallback.
"""
from .model import (BRANCH, NODE, DLWGD as DLWGDModel, ModelNode, closestnode, isawgd, isleaf, isroot, issp,
                    iswgd, iswgdafter, iswgm, iswgmafter, iswgt, iswgtafter, nonwgdchild, nonwgdparent,
                    postwalk, prewalk, update_)
from .newick import readnw
from .parallel import allreduce_sum_, dist_info, shard_range
from .priors import BMRevJumpPrior, CRRevJumpPrior, IRRevJumpPrior
from .profile import (DLWGD_ as DLWGD, PArray, extend_, gradient, gradient_cr, logpdf_, logpdfroot, profile, rev_, set_,
                      shrink_)


def addwgd_(model, n, t, q):
    """addwgd!(d, n, t, q)  (src/model.jl:106-114)"""
    return model.addwgd_(n, t, q)


def addwgt_(model, n, t, q):
    """addwgt!(d, n, t, q)  (src/model.jl:116-124)"""
    return model.addwgt_(n, t, q)


def removewgd_(model, n, reindex=True, set=True):
    """removewgd!(d, n, reindex, set)  (src/model.jl:158-175)"""
    return model.removewgd_(n, reindex, set)


def asvector(model):
    return model.asvector()

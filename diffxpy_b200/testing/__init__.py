from .tests import wald, lrt, wald_repeated, t_test, rank_test, two_sample, pairwise, versus_rest, partition, continuous_1d, _fit
from .det import DifferentialExpressionTestWald, DifferentialExpressionTestLRT, DifferentialExpressionTestTT, \
    DifferentialExpressionTestRank
from . import utils, correction
from .det_pair import DifferentialExpressionTestZTest, DifferentialExpressionTestZTestLazy, \
    DifferentialExpressionTestPairwiseStandard, DifferentialExpressionTestVsRest, DifferentialExpressionTestByPartition
from .det_cont import DifferentialExpressionTestWaldCont, DifferentialExpressionTestLRTCont

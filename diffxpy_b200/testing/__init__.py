from .tests import wald, lrt, wald_repeated, t_test, rank_test, two_sample, _fit
from .det import DifferentialExpressionTestWald, DifferentialExpressionTestLRT, DifferentialExpressionTestTT, \
    DifferentialExpressionTestRank
from . import utils, correction

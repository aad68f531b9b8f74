def ifelse(cond, a, b):
    raise NotImplementedError('ifelse is only used by the L-BFGS branch, which the stand-in does not cover')

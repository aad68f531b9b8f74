def hessian(*a, **k):
    raise NotImplementedError('no autodiff in the aesara stand-in')


jacobian = hessian

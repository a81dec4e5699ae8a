import numpy as np
from scipy.special import erf, erfc, log1p
from scipy.optimize import least_squares
x = np.linspace(1e-3, 6.0, 6000)
Phi = 0.5*erfc(-x/np.sqrt(2)); Q = 0.5*erfc(x/np.sqrt(2))
u = 0.5*(np.log(Phi)-np.log(Q))
phi = np.exp(-x*x/2)/np.sqrt(2*np.pi)
def model(c, x):
    x2 = x*x
    p = c[-1]
    for ci in c[-2::-1]: p = p*x2 + ci
    return x*p
def gel(c, x):
    t = np.tanh(model(c, x)); return 0.5*(1+t)
for deg in (2,3,4):
    def res(c):
        # error in Phi(-x) weighted by x (the gelu abs error on the negative side is x*err) and at +x too
        P = gel(c, x)
        return np.concatenate([(P-Phi)*x, ((1-P)-Q)*x*30])
    c0 = np.zeros(deg+1); c0[0]=0.7978845608; c0[1]=0.7978845608*0.044715
    r = least_squares(res, c0, xtol=1e-15, ftol=1e-15, gtol=1e-15)
    c = r.x
    P = gel(c, x)
    # derivative of approx: hp + 0.5 x (1-t^2) u'(x)
    x2=x*x
    dp = np.polyval((c*(2*np.arange(len(c))+1))[::-1], x2)
    t = np.tanh(model(c,x))
    d_pos = P + 0.5*x*(1-t*t)*dp
    d_exact = Phi + x*phi
    d_neg = (1-P) - 0.5*x*(1-t*t)*dp ; d_neg_exact = Q - x*phi
    print(deg, list(c), "max|gelu err| +x", np.abs(x*(P-Phi)).max(), " -x", np.abs(x*((1-P)-Q)).max(), "rel -x (x<4)", (np.abs(((1-P)-Q)/Q))[x<4].max(),
          "deriv err", np.abs(d_pos-d_exact).max(), np.abs(d_neg-d_neg_exact).max())

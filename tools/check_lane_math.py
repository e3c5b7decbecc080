"""Numpy restatement of the branch-free exp / log / log1p of csrc/dx_fit_lane.cu (same range reduction, same coefficients,
separate multiply-add instead of fma): checks them against numpy on random arguments of the ranges the kernel feeds them."""
import numpy as np, struct
from math import factorial
def my_exp(x):
    L2E=1.4426950408889634; ln2_hi=6.93147180369123816490e-01; ln2_lo=1.90821492927058770002e-10
    n=np.rint(x*L2E)
    r=x-n*ln2_hi; r=r-n*ln2_lo
    p=np.full_like(x,1.0/factorial(13))
    for k in range(12,-1,-1): p=p*r+1.0/factorial(k)
    return np.ldexp(p,n.astype(np.int64))
def my_log(x):
    m,e=np.frexp(x)   # m in [0.5,1)
    m=m*2; e=e-1       # [1,2)
    big=m>1.4142135623730951
    m=np.where(big,m*0.5,m); e=np.where(big,e+1,e).astype(np.float64)
    f=m-1.0
    s=f/(2.0+f)
    z=s*s; w=z*z
    Lg1=6.666666666666735130e-01;Lg2=3.999999999940941908e-01;Lg3=2.857142874366239149e-01;Lg4=2.222219843214978396e-01
    Lg5=1.818357216161805012e-01;Lg6=1.531383769920937332e-01;Lg7=1.479819860511658591e-01
    t1=w*(Lg2+w*(Lg4+w*Lg6)); t2=z*(Lg1+w*(Lg3+w*(Lg5+w*Lg7)))
    R=t2+t1
    hfsq=0.5*f*f
    ln2_hi=6.93147180369123816490e-01; ln2_lo=1.90821492927058770002e-10
    return e*ln2_hi-((hfsq-(s*(hfsq+R)+e*ln2_lo))-f)
def my_log1p(x):
    u=1.0+x
    c=(x-(u-1.0))/u
    return my_log(u)+c
rng=np.random.default_rng(0)
x=rng.uniform(-298,284,2000000)
ref=np.exp(x); got=my_exp(x)
print("exp max rel err (ulp)", np.max(np.abs(got-ref)/np.spacing(ref)))
x=np.exp(rng.uniform(-290,290,2000000)); x=np.concatenate([x, rng.uniform(0.5,2.0,1000000), 1+rng.uniform(-1e-6,1e-6,100000)])
ref=np.log(x); got=my_log(x)
print("log max err (ulp)", np.max(np.abs(got-ref)/np.spacing(np.abs(ref)+1e-300)), "max abs near 1", np.max(np.abs(got-ref)[-100000:]))
x=np.exp(rng.uniform(-60,60,3000000))
ref=np.log1p(x); got=my_log1p(x)
print("log1p max err (ulp)", np.max(np.abs(got-ref)/np.spacing(ref)))

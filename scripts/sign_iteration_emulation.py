"""CPU emulation (numpy) of the large-block sign iteration of csrc/dense_proj.cu: polynomial steps needed with the current
spectral bound ||A0^2||_F^(1/2), with the sharper ||A0^4||_F^(1/4) that falls out of the second product of step 0, and with the
exact 2-norm — for random symmetric matrices and for ADMM-like complementary spectra.  Result: 9 steps (29
products) is the floor; only random matrices of n >= 800 gain a step from a sharper bound, the solver's iterates do not."""
import numpy as np, sys
P=[(4.2567538552083883, -12.637529108386531, 9.3807751804785546),(4.2554599297572135, -12.626635057799284, 9.3711750746066897),(4.249950912774862, -12.580323192655207, 9.3303722673722636),(4.226491089776534, -12.384384556067777, 9.1578934252368036),(4.1267873247419073, -11.574501430286471, 8.4477140393475434),(3.7225306718083893, -8.6538466523498734, 5.9313159658509376),(2.658148855151357, -3.3824715696999976, 1.7243227088946445)]
NS=(15/8,-10/8,3/8)
def run(A, sharp):
    n=A.shape[0]
    A0=A/np.linalg.norm(A)
    X2=A0@A0
    if sharp==0: sc=np.linalg.norm(X2)**-0.5
    elif sharp==1: sc=np.linalg.norm(X2@X2)**-0.25
    else: sc=1/np.linalg.norm(A0,2)
    X=sc*A0
    res=[]; steps=0
    for k in range(40):
        a,b,c=P[k] if k<7 else NS
        X2=X@X
        r=np.linalg.norm(X2-np.eye(n))**2
        res.append(r)
        if k>=1:
            if r<1e-10: break
            if k>=8 and len(res)>=2 and abs(res[-2]-r)<=1e-13*r: break
        Z=c*(X2@X2)+b*X2
        X=X@Z+a*X
        steps+=1
    return steps, X
rng=np.random.default_rng(0)
for n in [200,800,2000]:
    for trial in range(3):
        M=rng.standard_normal((n,n)); A=(M+M.T)/2
        out=[]
        for sharp in (0,1,2):
            s,X=run(A,sharp); out.append(s)
        print(n,trial,'steps old/4th-power/exact-norm',out, 'products', [3*s+2 for s in out])
# ADMM-like: low-rank plus/minus structure (X - sigma S): half positive rank r, rest negative
for n in [800]:
    Q,_=np.linalg.qr(rng.standard_normal((n,n)))
    lam=np.concatenate([rng.uniform(0.5,2,n//3), -rng.uniform(0.5,2,n-n//3)])
    A=(Q*lam)@Q.T
    print('complementary', n, [run(A,s)[0] for s in (0,1,2)])

"""Model of the two-digit base-n engine, T=2 lanes x 32 limbs per digit (|n| = 2048 exactly)."""
import random
LB=32; D=64; M32=(1<<32)-1; W=1<<(32*D); BK=1<<(32*LB)
def limbs(x,n): return [(x>>(32*i))&M32 for i in range(n)]
def val(l): return sum(v<<(32*i) for i,v in enumerate(l))
def split(x): return [x & (BK-1), x>>(32*LB)]          # digit -> 2 blocks (lane0, lane1)

def blockmul_cols(a, b, clo, chi):
    """own 32-limb block a x 64-limb b, product-scanning columns [clo, chi), 3-word accumulator starting at 0.
    Returns dict col->limb plus final carry words (as the 2 limbs after chi-1)."""
    acc=0; out={}
    for c in range(clo, chi):
        for i in range(LB):
            j=c-i
            if 0<=j<len(b): acc+=a[i]*b[j]
        out[c]=acc&M32; acc>>=32
    out[chi]=acc&M32; out[chi+1]=(acc>>32)&M32
    assert acc>>64==0
    return out

def full_product(A,B):
    """A,B digits (ints <W). lane g: R_g = A_g x B (96 cols). returns (Plo, Phi) digits via slice assembly."""
    Ab=split(A); Bl=limbs(B,D)
    R=[blockmul_cols(limbs(Ab[g],LB),Bl,0,95) for g in range(2)]   # cols 0..94 + carry limb 95 (,96=0)
    for g in range(2): assert R[g][96]==0
    Rg=[[R[g][c] for c in range(96)] for g in range(2)]
    # slices
    sA=[None,None]; cA=[0,0]; sB=[None,None]; cB=[0,0]
    for g in range(2):
        c=0; s=[]
        for t in range(LB):
            v=Rg[0][32*g+t]+(Rg[1][t] if g else 0)+c; s.append(v&M32); c=v>>32
        sA[g]=s; cA[g]=c
        c=0; s=[]
        for t in range(LB):
            v=Rg[g][64+t]+(0 if g else Rg[1][32+t])+c; s.append(v&M32); c=v>>32
        sB[g]=s; cB[g]=c
    assert cA[0]==0
    # carry fixups: cA[1] -> lane0 sliceB ; then cB[0](updated) -> lane1 sliceB
    v=val(sB[0])+cA[1]; cB0=cB[0]+(v>>(32*LB)); sB[0]=limbs(v&(BK-1),LB)
    v=val(sB[1])+cB0; assert v>>(32*LB)==0 and cB[1]==0; sB[1]=limbs(v,LB)
    Plo=val(sA[0])+(val(sA[1])<<(32*LB)); Phi=val(sB[0])+(val(sB[1])<<(32*LB))
    assert Plo+Phi*W==A*B
    return Plo,Phi

def barrett(Plo,Phi,n,mu1, stats):
    """returns (q, r) exact with P = Phi*W+Plo < n*W"""
    Hb=split(Phi); Ml=limbs(mu1,D)
    # lane g: own H_g x mu' columns >= 30 (local) .. 95
    R=[blockmul_cols(limbs(Hb[g],LB),Ml,30,95) for g in range(2)]
    # global limb = 32g + c ; want floor(sum/W): limbs>=64, with guards
    tot=0
    for g in range(2):
        for c,v in R[g].items():
            tot+=v<<(32*(32*g+c))
    that=tot>>(32*D)
    t=(Phi*mu1)>>(32*D)
    assert 0<=t-that<=1,(t-that)
    qh=Phi+that
    P=Phi*W+Plo
    q=P//n
    assert 0<=q-qh<=3,(q-qh)
    stats[q-qh]=stats.get(q-qh,0)+1
    # low product columns <= 64 of qh*n (65 limbs)
    Qb=split(qh); Nl=limbs(n,D)
    assert qh<W
    low=0
    for g in range(2):
        Rl=blockmul_cols(limbs(Qb[g],LB),Nl,0,65)   # cols 0..64 (+2 carry limbs ignored beyond)
        for c in range(65):
            gl=32*g+c
            if gl<=64: low+=Rl[c]<<(32*gl)
    MOD=1<<(32*65)
    low%=MOD
    assert low==(qh*n)%MOD
    r=((P%MOD)-low)%MOD
    k=0
    while r>=n:
        r-=n; qh+=1; k+=1
    assert k<=3 and qh==q and r==P%n
    return qh,r

def sqr2(X0,X1,n,mu1,stats):
    Plo,Phi=full_product(X0,X0)
    Q,R=barrett(Plo,Phi,n,mu1,stats)
    Ulo,Uhi=full_product(X0,X1)
    _,U=barrett(Ulo,Uhi,n,mu1,stats)
    X1n=2*U+Q
    for _ in range(2):
        if X1n>=n: X1n-=n
    assert X1n<n
    return R,X1n
def mul2(X0,X1,Y0,Y1,n,mu1,stats):
    Plo,Phi=full_product(X0,Y0)
    Q,R=barrett(Plo,Phi,n,mu1,stats)
    a,b=full_product(X0,Y1); _,U=barrett(a,b,n,mu1,stats)
    a,b=full_product(X1,Y0); _,V=barrett(a,b,n,mu1,stats)
    X1n=U+V+Q
    for _ in range(2):
        if X1n>=n: X1n-=n
    assert X1n<n
    return R,X1n

random.seed(3)
stats={}
for trial in range(6):
    n=random.getrandbits(2048)|(1<<2047)|1
    if trial==1: n=(1<<2048)-1
    if trial==2: n=(1<<2047)+1
    mu=(W*W)//n; mu1=mu-W; assert 0<mu1<W,(mu1.bit_length())
    nn=n*n
    x=random.randrange(nn); y=random.randrange(nn)
    X0,X1=x%n,x//n; Y0,Y1=y%n,y//n
    for it in range(4):
        X0,X1=sqr2(X0,X1,n,mu1,stats); x=x*x%nn
        assert X0+X1*n==x
        X0,X1=mul2(X0,X1,Y0,Y1,n,mu1,stats); x=x*y%nn
        assert X0+X1*n==x
    # extremes
    for (a0,a1) in [(n-1,n-1),(0,n-1),(n-1,0),(1,0)]:
        r0,r1=sqr2(a0,a1,n,mu1,stats); assert r0+r1*n==pow(a0+a1*n,2,nn)
print("ok",stats)

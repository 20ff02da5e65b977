import random
M32=0xffffffff
def model_montmul(a,b,n,T,L,init=0,a2=None,b2=None):
    # a2, b2: a second product row per step (cios_step2: acc += a*b_i + a2*b2_i + q*n), as the fused two-digit multiply uses
    S=T*L
    R=1<<(32*S)
    n0inv=(-pow(n,-1,1<<32))&M32
    def limbs(x): return [[(x>>(32*(g*L+j)))&M32 for j in range(L)] for g in range(T)]
    A=limbs(a);B=limbs(b);N=limbs(n)
    A2=limbs(a2) if a2 is not None else None; B2=limbs(b2) if a2 is not None else None
    # per lane arrays as integers: X = list of L+2 limbs
    E=[[(init>>(32*(g*L+k)))&M32 if (k<L or g==T-1) else 0 for k in range(L+2)] for g in range(T)]
    assert init>>(32*(S+2))==0
    O=[[0]*(L+2) for _ in range(T)]
    def val(arr,lo,cnt): # little endian to int
        v=0
        for k in range(cnt): v|=arr[lo+k]<<(32*k)
        return v
    def put(arr,lo,cnt,v):
        for k in range(cnt):
            arr[lo+k]=(v>>(32*k))&M32
        return v>>(32*cnt)
    def mad_even(X,a,bb):
        # chain over X[0..L+1]
        v=val(X,0,L+2)
        for j in range(0,L,2): v+= (a[j]*bb)<<(32*j)
        ov=put(X,0,L+2,v); assert ov==0,"even overflow"
    def mad_odd(Z,a,bb):
        v=val(Z,0,L+2)
        for j in range(0,L,2): v+= (a[j+1]*bb)<<(32*j)
        ov=put(Z,0,L+2,v); assert ov==0,"odd overflow"
    def step(X,Y,bvals,b2v=None):
        # X,Y: per-lane arrays
        ins=[Y[g+1][0] if g<T-1 else 0 for g in range(T)]
        Zs=[]
        for g in range(T):
            y=Y[g]; x=X[g]
            v=val(y,L,2)+ins[g]; ov=put(y,L,2,v); assert ov==0
            s=x[0]+y[1]; x[0]=s&M32; c=s>>32
            Z=[0]*(L+2)
            for j in range(0,L,2):
                t=A[g][j+1]*bvals[g]+val(y,j+2,2)+c
                Z[j]=t&M32; Z[j+1]=(t>>32)&M32; c=t>>64
            Z[L]=c; Z[L+1]=0
            if b2v is not None: mad_odd(Z,A2[g],b2v)
            mad_even(x,A[g],bvals[g])
            if b2v is not None: mad_even(x,A2[g],b2v)
            Zs.append(Z)
        q=(X[0][0]*n0inv)&M32
        for g in range(T):
            mad_even(X[g],N[g],q)
            mad_odd(Zs[g],N[g],q)
            Y[g][:]=Zs[g]
    for owner in range(T):
        for j in range(0,L,2):
            b0=B[owner][j]; b1=B[owner][j+1]
            if a2 is None:
                step(E,O,[b0]*T)
                step(O,E,[b1]*T)
            else:
                step(E,O,[b0]*T,B2[owner][j])
                step(O,E,[b1]*T,B2[owner][j+1])
    # merge
    tot=0
    res=[]
    for g in range(T):
        inn=O[g+1][0] if g<T-1 else 0
        v=val(O[g],L,2)+inn; assert put(O[g],L,2,v)==0
        v=val(E[g],0,L+2)+val(O[g],1,L+1)
        assert put(E[g],0,L+2,v)==0
        tot+=val(E[g],0,L+2)<<(32*g*L)
    assert E[0] is not None
    # lane0's O[0] must be zero? (limb -1)
    ab=a*b+(a2*b2 if a2 is not None else 0)
    exp=((init+ab)*pow(R,-1,n))%n
    assert tot%n==exp,(tot,exp)
    assert tot<(3 if a2 is not None else 2)*n+(2 if init else 0)
    assert tot*R-init-ab>=0 and (tot*R-init-ab)%n==0 and (tot*R-init-ab)//n<R
    # max top limbs
    return tot
def model_montmul_init(a,b,n,T,L,init,a2=None,b2=None):
    return model_montmul(a,b,n,T,L,init,a2,b2)
if __name__=='__main__':
  random.seed(1)
  for (T,L) in [(4,8),(8,8),(8,12),(8,16),(16,12),(16,16),(4,2),(32,2)]:
      S=T*L
      for it in range(20):
          n=random.getrandbits(32*S)|1|(1<<(32*S-1)) if it%2==0 else (random.getrandbits(32*S-random.randint(0,40))|1)
          if it==5: n=(1<<(32*S))-1
          a=random.randrange(n); b=random.randrange(n)
          if it==3: a=b=n-1
          model_montmul(a,b,n,T,L)
      print(T,L,"ok")

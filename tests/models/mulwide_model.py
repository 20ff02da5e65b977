import random
M32=0xffffffff
def mul_wide(a,b,T,L):
    S=T*L
    def limbs(x): return [[(x>>(32*(g*L+j)))&M32 for j in range(L)] for g in range(T)]
    A=limbs(a); Bl=[(b>>(32*i))&M32 for i in range(S)]
    E=[[0]*(L+2) for _ in range(T)]; O=[[0]*(L+2) for _ in range(T)]
    def val(arr,lo,cnt): return sum(arr[lo+k]<<(32*k) for k in range(cnt))
    def put(arr,lo,cnt,v):
        for k in range(cnt): arr[lo+k]=(v>>(32*k))&M32
        return v>>(32*cnt)
    def step(X,Y,bv):
        ins=[Y[g+1][0] if g<T-1 else 0 for g in range(T)]
        Zs=[]
        for g in range(T):
            y=Y[g]; x=X[g]
            assert put(y,L,2,val(y,L,2)+ins[g])==0
            s=x[0]+y[1]; x[0]=s&M32; c=s>>32
            Z=[0]*(L+2)
            for j in range(0,L,2):
                t=A[g][j+1]*bv+val(y,j+2,2)+c
                Z[j]=t&M32; Z[j+1]=(t>>32)&M32; c=t>>64
            Z[L]=c; Z[L+1]=0
            v=val(x,0,L+2)
            for j in range(0,L,2): v+=(A[g][j]*bv)<<(32*j)
            assert put(x,0,L+2,v)==0
            Zs.append(Z)
        for g in range(T): Y[g][:]=Zs[g]
    lo=[]
    X,Y=E,O
    for i in range(S):
        step(X,Y,Bl[i])
        lo.append(X[0][0])      # limb leaving lane 0
        X,Y=Y,X
    # after S steps (even): X is E again (roles alternate), finish as mont_mul tail with E=X, O=Y
    E2,O2=X,Y
    hi=0; ovs=[]
    for g in range(T):
        inn=O2[g+1][0] if g<T-1 else 0
        assert put(O2[g],L,2,val(O2[g],L,2)+inn)==0
        v=val(E2[g],0,L+2)+val(O2[g],1,L+1)
        assert put(E2[g],0,L+2,v)==0
        assert E2[g][L+1]==0
    tot=sum(val(E2[g],0,L+1)<<(32*g*L) for g in range(T))
    lo_v=sum(v<<(32*i) for i,v in enumerate(lo))
    assert lo_v+(tot<<(32*S))==a*b,(hex(a*b-lo_v-(tot<<(32*S))))
    assert E2[T-1][L]==0
random.seed(5)
for (T,L) in [(2,32),(4,16),(2,4),(8,2)]:
    S=T*L
    for it in range(12):
        a=random.getrandbits(32*S); b=random.getrandbits(32*S)
        if it==0: a=b=(1<<(32*S))-1
        if it==1: a=0
        mul_wide(a,b,T,L)
    print(T,L,"ok")

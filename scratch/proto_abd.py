import numpy as np, scipy.linalg as sla, sys
from scipy.linalg import lapack

def flame_rhs(nsp):
    q_heat = 3000.0*1e3*0.034; dT=600.0; Ts=600.0; l=3e-4; m0=34.2/1000; lam=0.07; p=2e6; tm=1500.0
    pe_q=0.0090168; d_ro=2.88e-4; pe_d=1.5e-3; ro_m=m0*p/(8.314*tm); A=1.3e5; E=5000*4.184; R=8.314
    qm=l**2/Ts; qs=l**2; m_reag=0.342
    nu=[1.0]+[-w for w in np.linspace(0.4,0.1,nsp-1)] if nsp>1 else [1.0]
    nv=2+2*nsp
    def f(y):
        T=y[0]*Ts+dT
        rate=A*np.exp(-E/(R*T))*y[2]*(ro_m/m_reag)
        out=np.zeros(nv)
        out[0]=y[1]/lam
        out[1]=y[1]*pe_q-q_heat*rate*qm
        for k in range(nsp):
            out[2+2*k]=y[3+2*k]/d_ro
            out[3+2*k]=y[3+2*k]*pe_d+nu[k]*rate*ro_m*qs
        return out
    def fy(y):
        T=y[0]*Ts+dT
        ex=A*np.exp(-E/(R*T))*(ro_m/m_reag)
        rate=ex*y[2]
        drdT=rate*E/(R*T*T)*Ts
        drdC=ex
        J=np.zeros((nv,nv))
        J[0,1]=1/lam
        J[1,1]=pe_q; J[1,0]=-q_heat*qm*drdT; J[1,2]=-q_heat*qm*drdC
        for k in range(nsp):
            J[2+2*k,3+2*k]=1/d_ro
            J[3+2*k,3+2*k]+=pe_d
            J[3+2*k,0]+=nu[k]*ro_m*qs*drdT
            J[3+2*k,2]+=nu[k]*ro_m*qs*drdC
        return J
    return nv,f,fy

def build(nv,f,fy,N,Y,h):
    A=np.zeros((N,nv,nv)); F=np.zeros((N,nv))
    for j in range(N):
        A[j]=-np.eye(nv)-h*fy(Y[j]); F[j]=Y[j+1]-Y[j]-h*f(Y[j])
    return A,F

def banded_solve(A,F,left,right):
    N,nv,_=A.shape
    # unknown positions
    full=(N+1)*nv
    isbc=np.zeros(full,bool)
    for v in left: isbc[v]=True
    for v in right: isbc[N*nv+v]=True
    pos=np.nonzero(~isbc)[0]; idx=-np.ones(full,int); idx[pos]=np.arange(len(pos))
    n=N*nv
    rows=[];cols=[];vals=[]
    for j in range(N):
        for i in range(nv):
            for c in range(nv):
                cc=idx[j*nv+c]
                if cc>=0 and A[j,i,c]!=0: rows.append(j*nv+i);cols.append(cc);vals.append(A[j,i,c])
            cc=idx[(j+1)*nv+i]
            if cc>=0: rows.append(j*nv+i);cols.append(cc);vals.append(1.0)
    rows=np.array(rows);cols=np.array(cols);vals=np.array(vals)
    kl=(rows-cols).max();ku=(cols-rows).max()
    ldab=2*kl+ku+1
    ab=np.zeros((ldab,n))
    ab[kl+ku+rows-cols,cols]=vals
    lu,piv,info=lapack.dgbtrf(ab,kl,ku)
    assert info==0,info
    x,info=lapack.dgbtrs(lu,kl,ku,F.reshape(-1).copy(),piv)
    dyfull=np.zeros(full); dyfull[pos]=x
    return dyfull.reshape(N+1,nv),(kl,ku)

def elim(M,W,w):
    """partial pivot GE on stacked M (2nv x nv), applied to W (2nv x m) and w"""
    M=M.copy();W=W.copy();w=w.copy()
    r,nv=M.shape
    for k in range(nv):
        p=k+np.argmax(np.abs(M[k:,k]))
        if p!=k:
            M[[k,p]]=M[[p,k]];W[[k,p]]=W[[p,k]];w[[k,p]]=w[[p,k]]
        l=M[k+1:,k]/M[k,k]
        M[k+1:,k:]-=np.outer(l,M[k,k:]); W[k+1:]-=np.outer(l,W[k]); w[k+1:]-=l*w[k]
    return M,W,w

def condense(As,Bs,rs):
    """sequential compaction of rows As[i] Y_i + Bs[i] Y_{i+1} = rs[i]; returns C,D,g and backsub info"""
    nv=As[0].shape[0]
    C=As[0].copy();D=Bs[0].copy();g=rs[0].copy()
    info=[]
    Z=np.zeros((nv,nv))
    for i in range(1,len(As)):
        M=np.vstack([D,As[i]]); W=np.block([[C,Z],[Z,Bs[i]]]); w=np.concatenate([g,rs[i]])
        M,W,w=elim(M,W,w)
        U=M[:nv]; E=W[:nv,:nv]; Fm=W[:nv,nv:]; e=w[:nv]
        info.append((U,E,Fm,e))
        C=W[nv:,:nv];D=W[nv:,nv:];g=w[nv:]
    return C,D,g,info

def backsub(info,Ys,Ye):
    """given slab end nodes, recover interior nodes (list index 1..len-1)"""
    out=[None]*len(info)
    Yn=Ye
    for i in range(len(info)-1,-1,-1):
        U,E,Fm,e=info[i]
        Yi=np.linalg.solve(np.triu(U),e-E@Ys-Fm@Yn)
        out[i]=Yi;Yn=Yi
    return out

def abd_solve(A,F,left,right,S):
    N,nv,_=A.shape
    I=np.eye(nv)
    levels=[]
    As=[A[j] for j in range(N)];Bs=[I]*N;rs=[F[j] for j in range(N)]
    while len(As)>1:
        nA=[];nB=[];nr=[];infos=[]
        for s in range(0,len(As),S):
            C,D,g,info=condense(As[s:s+S],Bs[s:s+S],rs[s:s+S])
            nA.append(C);nB.append(D);nr.append(g);infos.append(info)
        levels.append((infos,S,len(As)))
        As,Bs,rs=nA,nB,nr
    C,D,g=As[0],Bs[0],rs[0]
    # top: C dY0 + D dYN = g with BC
    lfree=[v for v in range(nv) if v not in left]; rfree=[v for v in range(nv) if v not in right]
    K=np.hstack([C[:,lfree],D[:,rfree]])
    z=np.linalg.solve(K,g)
    Y0=np.zeros(nv);YN=np.zeros(nv);Y0[lfree]=z[:len(lfree)];YN[rfree]=z[len(lfree):]
    nodes={0:Y0}
    # node values at current level: list of boundary nodes
    vals=[Y0,YN]  # nodes of the top level (1 row -> 2 nodes)
    for infos,S,nrows in reversed(levels):
        newvals=[None]*(nrows+1)
        for si,info in enumerate(infos):
            s=si*S
            Ys=vals[si];Ye=vals[si+1]
            newvals[s]=Ys
            e=min(s+S,nrows)
            newvals[e]=Ye
            inner=backsub(info,Ys,Ye)
            for i,Yi in enumerate(inner): newvals[s+1+i]=Yi
        vals=newvals
    return np.array(vals)

if __name__=="__main__":
    nsp=int(sys.argv[1]) if len(sys.argv)>1 else 5
    N=int(sys.argv[2]) if len(sys.argv)>2 else 2000
    S=int(sys.argv[3]) if len(sys.argv)>3 else 32
    nv,f,fy=flame_rhs(nsp)
    h=1.0/N
    left=[0]+[2+2*k for k in range(nsp)]; right=[1]+[3+2*k for k in range(nsp)]
    Y=np.full((N+1,nv),0.99)
    Y[0,0]=(1000-600)/600; Y[0,2]=1.0
    for k in range(1,nsp): Y[0,2+2*k]=1e-3
    Y[N,1]=1e-10
    for k in range(nsp): Y[N,3+2*k]=1e-7
    for it in range(6):
        A,F=build(nv,f,fy,N,Y,h)
        d1,(kl,ku)=banded_solve(A,F,left,right)
        d2=abd_solve(A,F,left,right,S)
        rel=np.linalg.norm(d1-d2)/np.linalg.norm(d1)
        relinf=np.abs(d1-d2).max()/np.abs(d1).max()
        print(it,"kl,ku",kl,ku,"|dy|",np.linalg.norm(d1),"rel2",rel,"relinf",relinf)
        Y=Y-d1

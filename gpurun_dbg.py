import os, sys
os.environ["B200RL_TC"]="1"
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
import devlib
from test_op_kernels import _problem,_run,_seg
dev=devlib.Dev("cuda")
M=N=256; K=256
dZ=np.eye(M,N,dtype=np.float32)
W=(np.arange(N)[:,None]*1000+np.arange(K)[None,:]).astype(np.float32)   # W[n][k] = 1000 n + k
dX=dev.zeros((M,K))
p=_problem(M,K,dX,K); _seg(p,dev.to_dev(dZ),N,dev.to_dev(W),K,N)
_run(dev,1,[p])
out=dev.from_dev(dX,(M,K))
print("dgrad identity: expect out[m][k]=1000m+k")
np.set_printoptions(linewidth=200, suppress=True)
print(out[:4,:12]); print(out[8:10,:12]); print(out[33,:8], out[130,:8], out[255,250:])
print("nonzero frac", (out!=0).mean(), "match frac", (out==W).mean())
# wgrad: dW[n][k] = sum_m dZ[m][n] X[m][k]; dZ = eye -> dW = X
X=(np.arange(M)[:,None]*1000+np.arange(K)[None,:]).astype(np.float32)
dW=dev.zeros((N,K))
q=_problem(N,K,dW,K); _seg(q,dev.to_dev(dZ),N,dev.to_dev(X),K,M)
_run(dev,2,[q])
o2=dev.from_dev(dW,(N,K))
print("wgrad identity: expect", X[:2,:6]); print(o2[:4,:12]); print("nonzero frac",(o2!=0).mean(),"match",(o2==X).mean())

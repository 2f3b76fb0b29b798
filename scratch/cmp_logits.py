import numpy as np, sys
a=np.fromfile(sys.argv[1],np.float32); b=np.fromfile(sys.argv[2],np.float32); n=int(sys.argv[3])
a=a.reshape(n,-1); b=b.reshape(n,-1)
for i in range(n): print(i, "maxdiff %.3e"%np.abs(a[i]-b[i]).max(), "scale %.3f"%np.abs(b[i]).max(), "argmax", a[i].argmax(), b[i].argmax())

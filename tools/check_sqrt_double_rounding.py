"""(float)sqrt((double)x) == sqrtf(x) for EVERY non-negative binary32 x (2,139,095,041 bit patterns, 0 mismatches, ~20 s):
the licence for pgm::sqrt_f64_as_f32 (crux_b200/csrc/pg_math.cuh).  numpy float32 sqrt is the correctly rounded IEEE root."""
import numpy as np, time
t=time.time(); bad=0; N=1<<26
for start in range(0, 0x7f800000+1, N):   # every non-negative float32 bit pattern up to +inf
    end=min(start+N, 0x7f800000+1)
    x=np.arange(start,end,dtype=np.uint32).view(np.float32)
    a=np.sqrt(x)                          # correctly rounded binary32 root
    b=np.sqrt(x.astype(np.float64)).astype(np.float32)
    bad+=int((a.view(np.uint32)!=b.view(np.uint32)).sum())
print("patterns checked", 0x7f800000+1, "mismatches", bad, "seconds", round(time.time()-t,1))

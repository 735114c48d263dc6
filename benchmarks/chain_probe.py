import sys, numpy as np
sys.path.insert(0,'/root/repo')
import sqaod_b200 as sq
rng=np.random.default_rng(1)
for N,m in ((128,128),(512,512),(1024,128),(2048,256),(8192,512)):
    A=rng.random((N,N),dtype=np.float32)-np.float32(0.5); W=np.triu(A)+np.triu(A,1).T
    ann=sq.dense_graph_annealer(W,sq.minimize,np.float32,n_trotters=m); ann.seed(1); ann.prepare(); ann.randomize_spin()
    for _ in range(3): ann.anneal_one_step(0.01,50.0)
    s0=ann.get_stats(); 
    import time; ann._device.synchronize(); t=time.perf_counter()
    n=10
    for _ in range(n): ann.anneal_one_step(0.01,50.0)
    ann._device.synchronize(); dt=(time.perf_counter()-t)/n*1e3
    s1=ann.get_stats(); G=min(148,m)
    print('N=%d m=%d: %.3f ms/step; per CTA busy: dot %.3f ms chain %.3f ms' % (N,m,dt,(s1['barrier_cycles_dot']-s0['barrier_cycles_dot'])/G/1.965e6/n,(s1['barrier_cycles_chain']-s0['barrier_cycles_chain'])/G/1.965e6/n))

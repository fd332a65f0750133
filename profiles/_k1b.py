import sys, os, torch, numpy as np, subprocess
if len(sys.argv) > 1:
    sys.path.insert(0,'.')
    import taxidispatcher_b200 as td
    from taxidispatcher_b200.dispatch import _ptr, _stream
    from oracle import gen_inputs as g
    eng=td.engine(); lib=eng.lib
    cab_to, cust_from = g.config5b()
    d=torch.from_numpy(g.stand_distances(4000)).cuda(); cab=torch.from_numpy(cab_to).cuda(); cu=torch.from_numpy(cust_from).cuda()
    out=torch.empty((20000,20000),dtype=torch.int32,device='cuda')
    flush=torch.empty(256<<20,dtype=torch.uint8,device='cuda')
    def t(fn,reps=7):
        for _ in range(3): fn()
        torch.cuda.synchronize(); ms=[]
        for _ in range(reps):
            flush.fill_(1)
            a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
        return sorted(ms)[len(ms)//2]
    def plain(): lib.td_cost_matrix_rows(_ptr(d),4000,_ptr(cab),20000,_ptr(cu),20000,250000,-1,0,20000,_ptr(out),None,0,_stream())
    print(sys.argv[1], 'plain ms %.4f' % t(plain))
else:
    for store in (0,1,2):
        for per_sm in (1,2,4,6):
            env=dict(os.environ, TD_K1_STORE=str(store), TD_K1_PER_SM=str(per_sm))
            subprocess.run([sys.executable, __file__, 'store=%d per_sm=%d'%(store,per_sm)], env=env)

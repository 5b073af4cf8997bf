import torch, time, threading, numpy as np
B,H,W=64,1080,1920
def t_copy(t, n=5):
    torch.cuda.synchronize()
    t0=time.time()
    for _ in range(n):
        d=t.to('cuda', non_blocking=True)
    t1=time.time()
    torch.cuda.synchronize()
    t2=time.time()
    return (t1-t0)/n*1e3, (t2-t0)/n*1e3
a=torch.empty((B,H,W,3),dtype=torch.uint8)
t0=time.time(); p1=a.pin_memory(); print('pin_memory() s', time.time()-t0, p1.is_pinned())
t0=time.time(); p2=torch.empty((B,H,W,3),dtype=torch.uint8,pin_memory=True); print('empty(pin) s', time.time()-t0, p2.is_pinned())
print('pageable  issue/total ms', t_copy(a))
print('pin_memory() issue/total ms', t_copy(p1))
print('view of pinned', p1[:60].is_pinned(), t_copy(p1[:60]))
print('empty(pin) issue/total ms', t_copy(p2))
src=[np.random.randint(0,255,(H,W,3),dtype=np.uint8) for _ in range(64)]
stop=False
def worker():
    arr=p2.numpy()
    while not stop:
        for i in range(64): arr[i]=src[i]
th=threading.Thread(target=worker,daemon=True); th.start()
time.sleep(0.2)
print('with a copier thread: issue/total ms', t_copy(p1))
stop=True

import torch, time
print("devices", torch.cuda.device_count())
print("peer 0->1", torch.cuda.can_device_access_peer(0,1))
a = torch.empty(1<<28, dtype=torch.float64, device="cuda:0")  # 2 GiB
b = torch.empty(1<<28, dtype=torch.float64, device="cuda:1")
for _ in range(2): b.copy_(a)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
t=time.time()
for _ in range(5): b.copy_(a)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
dt=(time.time()-t)/5
print("copy 0->1 GB/s", a.numel()*8/dt/1e9)

"""Developer tool: what the host pipeline can hope for -- H2D -> kernel -> D2H chunk pipelines in plain torch,
by chunk count and copies per chunk (PCIe 5 x16 on the B200 boxes: 55 GB/s each way, 97 GB/s both ways)."""
import os
import torch, time
n = 26214400
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
s_k = [torch.cuda.Stream() for _ in range(3)]
def pipe(chunks, pieces=1):
    k = n // chunks
    for i in range(chunks):
        sl = slice(i*k, (i+1)*k)
        with torch.cuda.stream(s_up):
            q = k // pieces
            for p in range(pieces): d_in[i*k+p*q:i*k+(p+1)*q].copy_(h_in[i*k+p*q:i*k+(p+1)*q], non_blocking=True)
            e1 = torch.cuda.Event(); e1.record()
        st = s_k[i % 3]
        with torch.cuda.stream(st):
            st.wait_event(e1)
            d_out[sl].copy_(d_in[sl])          # "kernel"
            e2 = torch.cuda.Event(); e2.record()
        with torch.cuda.stream(s_dn):
            s_dn.wait_event(e2)
            q = k // pieces
            for p in range(pieces): h_out[i*k+p*q:i*k+(p+1)*q].copy_(d_out[i*k+p*q:i*k+(p+1)*q], non_blocking=True)
def t(fn, reps=20):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
for c in (1, 2, 4, 8):
    print(c, "chunks:", round(t(lambda: pipe(c)), 3), "ms;  3 pieces per chunk:", round(t(lambda: pipe(c, 3)), 3))

import torch, time
n_in, n_out = 97_000_000, 86_000_000
h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both():
    h2d(); d2h()
def chunked(k=4):
    a, b = n_in // k, n_out // k
    for c in range(k):
        with torch.cuda.stream(s1): d_in[c*a:(c+1)*a].copy_(h_in[c*a:(c+1)*a], non_blocking=True)
        with torch.cuda.stream(s2): h_out[c*b:(c+1)*b].copy_(d_out[c*b:(c+1)*b], non_blocking=True)
print("h2d %.3f ms (%.1f GB/s)" % (t(h2d), n_in / t(h2d) / 1e6))
print("d2h %.3f ms (%.1f GB/s)" % (t(d2h), n_out / t(d2h) / 1e6))
print("both concurrently %.3f ms" % t(both))
print("both, 4 chunks each %.3f ms" % t(chunked))

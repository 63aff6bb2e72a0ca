"""Profiling driver: a few eager 7B decode steps (B=64, KV ring pre-filled to 2048) for ncu."""
import sys, torch
sys.path.insert(0, ".")
from rstnet_b200.lm import GPT, Config
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda", 0)
B, KV = 64, 2048
cfg = Config(block_size=4096, n_layer=layers, n_embd=4096, n_head=32, head_size=128, intermediate_size=11008, padded_vocab_size=152064,
             audio_card=2050, n_q=8, dep_q=8, codecformer_dim=1024, codecformer_heads=16, codecformer_layers=6,
             codecformer_dim_feedforward=4224, context=KV)
m = GPT(cfg, device=dev, dtype=torch.bfloat16).eval()
m.use_cuda_graphs = False
m.streaming_forever(B)
for kv in m._state.kv:
    kv.normal_()
m._state.offset.fill_(KV + 8)
seq = torch.randint(0, 2048, (B, 9, 1), device=dev)
for _ in range(steps):
    t = m.forward_step(seq)
torch.cuda.synchronize()
print("ok", t.shape)

"""ORACLE -- TEST INFRASTRUCTURE ONLY.  KV-cached restatement of the reference autoregressive decoder's
inference loop (reference models/ar_decoder.py:167-238): the reference re-runs prenet + positional encoding + the
whole 6-layer nn.TransformerDecoder on the growing prefix for every frame and keeps only the last position;
because the self-attention is causal, position t's activations never change once computed, so caching each
layer's self-attention keys / values (and projecting the encoder memory once) gives the same frames.

Layer arithmetic restated from torch.nn.TransformerDecoderLayer (post-norm, relu, batch_first; torch 2.11):
    x = LN1(x + SA(x));  x = LN2(x + MHA(x, memory));  x = LN3(x + W2 relu(W1 x + b1) + b2)
with nn.MultiheadAttention's packed in_proj ([q; k; v] rows), 1/sqrt(head_dim) scaling, no memory mask.
Pinned against the live reference by tests/golden/make_ar_decoder.py (frame-for-frame)."""
import math

import torch
import torch.nn.functional as F


def positional_encoding(max_len, d_model):
    # reference models/ar_decoder.py:303-312
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def decode(sd, hvar, n_layers, n_heads, max_len=None):
    """sd: state_dict of the reference PNCAARDecoder (torch tensors); hvar [B, Tfrm, d] -> mel [B, max_len, n_mels]."""
    B, Tfrm, d = hvar.shape
    max_len = Tfrm if max_len is None else max_len
    hd = d // n_heads
    n_mels = sd["mel_proj.weight"].shape[0]
    pe = sd["pos_encoding.pe"][0] if "pos_encoding.pe" in sd else positional_encoding(5000, d)
    L = lambda i, n: sd[f"decoder.layers.{i}.{n}"]
    # encoder memory projected once per layer
    mem_k, mem_v = [], []
    for i in range(n_layers):
        w, b = L(i, "multihead_attn.in_proj_weight"), L(i, "multihead_attn.in_proj_bias")
        mem_k.append(F.linear(hvar, w[d:2 * d], b[d:2 * d]).view(B, Tfrm, n_heads, hd))
        mem_v.append(F.linear(hvar, w[2 * d:], b[2 * d:]).view(B, Tfrm, n_heads, hd))
    k_cache = [hvar.new_zeros(B, max_len, n_heads, hd) for _ in range(n_layers)]
    v_cache = [hvar.new_zeros(B, max_len, n_heads, hd) for _ in range(n_layers)]
    frame = hvar.new_zeros(B, n_mels)                       # start token (reference :190)
    out = hvar.new_zeros(B, max_len, n_mels)
    scale = 1.0 / math.sqrt(hd)

    def attend(q, K, V):                                    # q [B, H, hd], K/V [B, T, H, hd]
        s = torch.einsum("bhd,bthd->bht", q, K) * scale
        return torch.einsum("bht,bthd->bhd", torch.softmax(s, dim=-1), V).reshape(B, d)

    for t in range(max_len):
        x = F.linear(F.relu(F.linear(frame, sd["prenet.0.weight"], sd["prenet.0.bias"])),
                     sd["prenet.3.weight"], sd["prenet.3.bias"]) + pe[t]
        for i in range(n_layers):
            qkv = F.linear(x, L(i, "self_attn.in_proj_weight"), L(i, "self_attn.in_proj_bias"))
            q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
            k_cache[i][:, t] = k.view(B, n_heads, hd)
            v_cache[i][:, t] = v.view(B, n_heads, hd)
            sa = attend(q.view(B, n_heads, hd), k_cache[i][:, : t + 1], v_cache[i][:, : t + 1])
            sa = F.linear(sa, L(i, "self_attn.out_proj.weight"), L(i, "self_attn.out_proj.bias"))
            x = F.layer_norm(x + sa, (d,), L(i, "norm1.weight"), L(i, "norm1.bias"))
            w, b = L(i, "multihead_attn.in_proj_weight"), L(i, "multihead_attn.in_proj_bias")
            q = F.linear(x, w[:d], b[:d]).view(B, n_heads, hd)
            ca = attend(q, mem_k[i], mem_v[i])
            ca = F.linear(ca, L(i, "multihead_attn.out_proj.weight"), L(i, "multihead_attn.out_proj.bias"))
            x = F.layer_norm(x + ca, (d,), L(i, "norm2.weight"), L(i, "norm2.bias"))
            ff = F.linear(F.relu(F.linear(x, L(i, "linear1.weight"), L(i, "linear1.bias"))),
                          L(i, "linear2.weight"), L(i, "linear2.bias"))
            x = F.layer_norm(x + ff, (d,), L(i, "norm3.weight"), L(i, "norm3.bias"))
        frame = F.linear(x, sd["mel_proj.weight"], sd["mel_proj.bias"])
        out[:, t] = frame
    return out

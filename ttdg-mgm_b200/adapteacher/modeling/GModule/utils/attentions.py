"""MultiHeadAttention / dot_attention; mirrors reference adapteacher/modeling/GModule/utils/attentions.py:25-116
('v2', the only version the matching head builds, mgm:463).  The attention map - the only output the hot path
uses (mgm:498) - comes from the sm_100a kernels; the (discarded) context branch is kept for API parity."""
import torch
import torch.nn as nn

from ttdg_b200 import ops


class dot_attention(nn.Module):
    def __init__(self, attention_dropout=0.0):
        super().__init__()
        self.dropout = nn.Dropout(attention_dropout)
        self.softmax = nn.Softmax(dim=2)


class MultiHeadAttention(nn.Module):
    def __init__(self, model_dim=256, num_heads=4, dropout=0.0, version='v2'):
        super().__init__()
        if version != 'v2' or num_heads != 1:
            raise NotImplementedError("the matching head uses MultiHeadAttention(256, 1, version='v2') (mgm:463)")
        self.dim_per_head = model_dim // num_heads
        self.num_heads = num_heads
        self.linear_k = nn.Linear(model_dim, self.dim_per_head * num_heads)
        self.linear_v = nn.Linear(model_dim, self.dim_per_head * num_heads)
        self.linear_q = nn.Linear(model_dim, self.dim_per_head * num_heads)
        self.dot_product_attention = dot_attention(dropout)
        self.linear_final = nn.Linear(model_dim, model_dim)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(model_dim)
        self.version = version
        self.philox_seed = 0x7464_6467           # production-mode dropout stream; advanced per call
        self.philox_offset = 0

    def adjacency(self, X, sizes, keep_masks=None):
        """Block-diagonal, zero-diagonal attention adjacency of all graphs stacked in X (mgm:496-502)."""
        p = self.dot_product_attention.dropout.p if self.training else 0.0
        A = ops.attention_adjacency(X, sizes, self.linear_q.weight, self.linear_q.bias, self.linear_k.weight,
                                    self.linear_k.bias, keep_masks=keep_masks, p_drop=p,
                                    seed=self.philox_seed, offset=self.philox_offset)
        if keep_masks is None and p > 0:
            self.philox_offset += int(X.shape[0]) ** 2
        return A

    def forward(self, key_value_query, attn_mask=None, need_output=True):
        if attn_mask is not None:
            raise NotImplementedError("attn_mask is never passed on the hot path")
        key, value, query = key_value_query
        if not (key is value and value is query):
            raise NotImplementedError("the matching head calls this with [x, x, x] (mgm:573)")
        # standalone callers get the raw (eval-mode) attention map; the hot path uses adjacency() (mgm:498-502)
        att = self._full_attention(key)
        out = None
        if need_output:
            with torch.no_grad():
                v = ops.linear(value, self.linear_v.weight, self.linear_v.bias)
                ctx = ops.gemm(att, v)
                o = ops.linear(ctx, self.linear_final.weight, self.linear_final.bias)
                out = torch.nn.functional.layer_norm(query + o, (query.shape[1],), self.layer_norm.weight, self.layer_norm.bias)
        return out, att

    def _full_attention(self, x):
        """Eval-mode softmax(q k^T / 16) including the diagonal (attentions.py:34-41)."""
        S = ops.attention_logits(x, self.linear_q.weight, self.linear_q.bias, self.linear_k.weight, self.linear_k.bias)
        return torch.softmax(S * float(self.dim_per_head ** -0.5), dim=1)

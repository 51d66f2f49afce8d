"""Word-level attention -- libmog edition of the reference's
``code/coco/attngan/GlobalAttention.py`` (same names and signatures).

``GlobalAttentionGeneral.forward`` = 1x1 conv on the word embeddings + ONE fused kernel for
bmm -> mask -> softmax -> bmm over NHWC pixels (``mog_word_attention_fwd/bwd``).  The
reference's mask handling tiles the (B,T) mask ``queryL`` times against a batch-major
(B*queryL, T) view (GlobalAttention.py:104-108), i.e. row (b*queryL+q) is masked with
``mask[(b*queryL+q) % B]``; that behaviour is reproduced by default (``cfg.MOG.MASK_QUIRK``).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .miscc.config import cfg


def conv1x1(in_planes, out_planes):
    from .model import Conv2d
    return Conv2d(in_planes, out_planes, 1, 1, 0, bias=False)


def func_attention(query, context, gamma1):
    """GlobalAttention.py:31-69 -- query: batch x ndf x queryL, context: batch x ndf x ih x iw ->
    (weightedContext batch x ndf x queryL, attn batch x queryL x ih x iw).  Forward only here; the
    training path uses the fused ``miscc.losses.words_loss``."""
    ctx = ops.nhwc(context)
    B, ih, iw, D = ctx.shape
    wei, attn = ops.func_attention_paired(query, ctx.reshape(B, ih * iw, D), gamma1)
    return wei, attn.reshape(B, -1, ih, iw)


class GlobalAttentionGeneral(nn.Module):
    def __init__(self, idf, cdf):
        super().__init__()
        self.conv_context = conv1x1(cdf, idf)
        self.sm = nn.Softmax(dim=1)  # kept for state/API parity; the softmax runs inside the fused kernel
        self.mask = None

    def applyMask(self, mask):
        self.mask = mask  # batch x sourceL

    def forward_nhwc(self, h, context):
        """h NHWC [B,ih,iw,idf]; context [B,cdf,T] -> (weighted context NHWC, attn [B,T,ih,iw])."""
        B, ih, iw, idf = h.shape
        T = context.shape[2]
        words = context.transpose(1, 2).contiguous()                     # [B,T,cdf]
        src = self.conv_context(words.reshape(B, T, 1, -1)).reshape(B, T, idf)
        out, attn = ops.word_attention(h.reshape(B, ih * iw, idf), src, self.mask,
                                       mask_quirk=cfg.MOG.MASK_QUIRK, want_attn=True)
        return out.reshape(B, ih, iw, idf), attn.reshape(B, T, ih, iw)

    def forward(self, input, context):
        """input: batch x idf x ih x iw (queryL=ihxiw); context: batch x cdf x sourceL (reference contract)."""
        out, attn = self.forward_nhwc(ops.nhwc(input), context)
        return ops.to_nchw_view(out), attn

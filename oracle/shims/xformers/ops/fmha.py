class BlockDiagonalMask:
    """Only what the reference uses: from_seqlens(q_seqlen, kv_seqlen)."""

    def __init__(self, q_seqlen, kv_seqlen):
        self.q_seqlen = [int(v) for v in q_seqlen]
        self.kv_seqlen = [int(v) for v in kv_seqlen]

    @classmethod
    def from_seqlens(cls, q_seqlen, kv_seqlen=None):
        return cls(q_seqlen, kv_seqlen if kv_seqlen is not None else q_seqlen)

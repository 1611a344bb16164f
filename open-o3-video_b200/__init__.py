"""open-o3-video_b200: B200-native (sm_100a) implementation of Open-o3-Video's RL
policy-objective hot path, behind the reference trainer's own call signatures.

    hidden states -> lm_head (tcgen05/TMEM GEMM fed by TMA) -> online log-softmax ->
    target gather -> GSPO ratio / clip / KL / loss + group advantages, fwd + bwd,
    plus the spatio-temporal reward numerics.

Everything numerical runs in `lib/libo3v.so` (CUDA, C ABI declared in include/o3v.h) and is
reached through ctypes (`_lib.py`); there is NO CPU or eager-PyTorch fallback: if the
library is missing or the device is not a B200 every op raises.

Public API (mirrors the reference, src/r1-v/src/open_r1/):
  logprob.per_token_logps / fused_logprob      trainer/grpo_trainer.py:371-384
  gspo.eos_mask / gspo_loss                    trainer/grpo_trainer.py:590-596, 635-706
  logprob.fused_logprob_gspo                   the whole step, chunked fwd+bwd in one call
  rewards.<reference reward names>             reward_func.py
  trainer.O3VB200TrainerMixin                  drop-in _get_per_token_logps / hot compute_loss
  sharded.*                                    vocab-parallel multi-GPU (NCCL / peer memory)
  ops (torch.ops.o3v.*)                        the same launches as registered operators (schema, fake
                                               kernels, autograd) for torch.compile / torch.export
"""
__version__ = "0.1.0"

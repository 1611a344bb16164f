"""CPU oracle for the Open-o3-Video GSPO policy-objective hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(``open-o3-video_b200/``) may import, call, link or execute anything in this
directory; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
(or as the timed CPU baseline), never as the thing shipped.

Parity status: the reference repository has no tests, golden vectors or
fixtures for this path.  The oracle is therefore pinned
against OUTPUTS OF THE REFERENCE ITSELF: ``tests/golden/gen_golden.py`` imports
the unmodified reference modules from ``/root/reference`` (through the stub
loader in ``oracle/ref_import.py``), runs them on seeded inputs and commits the
results under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every
oracle function against those vectors and against the known-answer values of
SURVEY.md Appendix B.  The inline loss block of ``compute_loss`` cannot be
imported (it needs a live model); it is restated line for line in
``oracle/gspo.py`` with every line cited, and its gradient is cross-checked
against torch autograd.

Modules
  logps.py    lm_head + log-softmax + gather   (grpo_trainer.py:371-384)
  gspo.py     EOS mask, KL, advantages, GSPO objective, metrics
              (grpo_trainer.py:590-596, 635-636, 658, 675-681, 691-706, 711, 737)
  rewards.py  numeric cores of the temporal / spatial rewards
              (reward_func.py:86-181, 184-236, 337-605)
  parse.py    completion text -> parsed rollout (the reference's regex / json / float extraction,
              reward_func.py:91-93, 119-126, 211-223, 308-335, 394-412, 437-449, 481-511) and the
              seeded text generators / mutators that the K6 tests use
  vstar.py    V-STAR scorer numerics (eval/test/eval_vstar.py:90-178)
  sft.py      SFT causal-LM cross-entropy (sft_multi_task.py:402-409)
  synth.py    seeded synthetic input generators shared by tests and bench
  ref_import.py  stub loader for the real reference (this container only)
"""

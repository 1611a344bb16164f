#!/usr/bin/env python
"""Level-2 integration: rewrite the three INLINE blocks of the reference's `compute_loss`
(src/r1-v/src/open_r1/trainer/grpo_trainer.py) into calls of `O3VB200TrainerMixin` methods, and make the class
inherit the mixin.  Inline code cannot be overridden by subclassing, so this is the one edit of the reference a
maintainer makes; everything else (vision prep, generate, decode, reward callables) stays the reference's code.

    python integration/patch_reference.py /path/to/grpo_trainer.py [--mode calls|fused] [--diff | -o OUT | --in-place]

The edit is located by identifiers, not by line numbers, and the script refuses to write unless every anchor is
found exactly once.  It ships as a script instead of a .patch file on purpose: a unified diff would carry ~80
lines of the reference's source into this repository; `--diff` prints that patch for review.

mode `calls` (default)                                               reference lines
  class Qwen2VLGRPOTrainer(Trainer)  ->  _ReferenceQwen2VLGRPOTrainer(Trainer), and at the end of the file
      class Qwen2VLGRPOTrainer(O3VB200TrainerMixin, _ReferenceQwen2VLGRPOTrainer)       :81
  EOS-mask block                     ->  self.o3v_completion_mask(...)       :590-596   K3a
  KL + advantages + objective + metrics -> self.compute_policy_loss(...)     :635-636, 658-738   K3
  (the two `_get_per_token_logps` calls already resolve to the mixin's: K1 on completion rows only)
mode `fused` additionally
  drops the policy `_get_per_token_logps` try/except (:611-617) and computes the policy pass, the loss and their
  backward in ONE call after the rewards are known:  self.o3v_fused_policy_loss(model, prompt_completion_ids,
  prompt_length, ref_per_token_logps, completion_mask, rewards_per_func, **prompt_inputs)   K1+K3+K2a+K2b
"""
import argparse
import difflib
import re
import sys

IMPORT_LINE = "from open_o3_video_b200.trainer import O3VB200TrainerMixin\n"


class PatchError(RuntimeError):
    pass


def _find(lines, pattern, start=0, what=None):
    rx = re.compile(pattern)
    hits = [i for i in range(start, len(lines)) if rx.search(lines[i])]
    if len(hits) != 1:
        raise PatchError("anchor %r: expected exactly one match, found %d" % (what or pattern, len(hits)))
    return hits[0]


def _indent(line):
    return line[: len(line) - len(line.lstrip())]


def patch_source(src: str, mode: str = "calls") -> str:
    if mode not in ("calls", "fused"):
        raise ValueError(mode)
    if "O3VB200TrainerMixin" in src:
        raise PatchError("already patched")
    lines = src.splitlines(keepends=True)

    # -- class statement: the reference class becomes the base, the exported name is the mixin laid over it (a
    #    mixin listed as a BASE of the reference class would lose to the class's own compute_loss /
    #    _get_per_token_logps in the MRO)
    c = _find(lines, r"^class\s+Qwen2VLGRPOTrainer\(Trainer\):", what="class statement")
    lines[c] = lines[c].replace("Qwen2VLGRPOTrainer(Trainer)", "_ReferenceQwen2VLGRPOTrainer(Trainer)")
    if not lines[-1].endswith("\n"):
        lines[-1] += "\n"
    lines += ["\n", "\n", IMPORT_LINE, "\n", "\n",
              "class Qwen2VLGRPOTrainer(O3VB200TrainerMixin, _ReferenceQwen2VLGRPOTrainer):\n",
              '    """The reference trainer with its hot path on the B200 kernels (integration/patch_reference.py)."""\n']

    body = _find(lines, r"^\s+def compute_loss\(self, model, inputs", what="compute_loss")

    # -- EOS mask block: from `is_eos = completion_ids == ...eos_token_id` to `completion_mask = (... <= eos_idx ...)`
    a = _find(lines, r"^\s+is_eos\s*=\s*completion_ids\s*==", body, "EOS block start")
    b = _find(lines, r"^\s+completion_mask\s*=.*sequence_indices\s*<=\s*eos_idx", body, "EOS block end")
    ind = _indent(lines[a])
    lines[a:b + 1] = [ind + "device = self.accelerator.device\n",
                      ind + "completion_mask = self.o3v_completion_mask(completion_ids)  # K3a (was :590-596)\n"]

    # -- KL lines
    k0 = _find(lines, r"^\s+x_clamped\s*=\s*torch\.clamp\(ref_per_token_logps\s*-\s*per_token_logps", body, "KL clamp")
    k1 = _find(lines, r"^\s+per_token_kl\s*=\s*torch\.exp\(x_clamped\)", body, "KL value")
    if k1 != k0 + 1:
        raise PatchError("KL lines are not adjacent")
    del lines[k0:k1 + 1]

    # -- advantages .. metrics: from `rewards = rewards_per_func.sum(dim=1)` to the line before `return loss`
    r0 = _find(lines, r"^\s+rewards\s*=\s*rewards_per_func\.sum\(dim=1\)", body, "reward sum")
    r1 = _find(lines, r"^\s+return loss\s*$", r0, "return loss")
    ind = _indent(lines[r0])
    if mode == "calls":
        call = [ind + "# K3: KL, group advantages, GSPO objective and the metrics in one launch (was :635-738)\n",
                ind + "loss = self.compute_policy_loss(per_token_logps, ref_per_token_logps, completion_mask, "
                      "rewards_per_func)\n", "\n"]
    else:
        call = [ind + "# K1 + K3 + K2a + K2b: policy pass, loss and their backward in one fused step (was :611-617, "
                      ":635-738)\n",
                ind + "loss = self.o3v_fused_policy_loss(model, prompt_completion_ids, prompt_length, "
                      "ref_per_token_logps, completion_mask, rewards_per_func, **prompt_inputs)\n", "\n"]
    lines[r0:r1] = call

    if mode == "fused":
        # the policy try/except: `try:` directly above `per_token_logps = self._get_per_token_logps(model, ...`
        # up to (not including) the `with torch.inference_mode():` of the reference pass
        p = _find(lines, r"^\s+per_token_logps\s*=\s*self\._get_per_token_logps\(model,\s*prompt_completion_ids,\s*\*\*",
                  body, "policy log-probs")
        if not re.match(r"^\s+try:\s*$", lines[p - 1]):
            raise PatchError("policy log-prob call is not inside the expected try block")
        q = _find(lines, r"^\s+with torch\.inference_mode\(\):\s*$", p, "reference pass")
        ind = _indent(lines[p - 1])
        lines[p - 1:q] = [ind + "# the policy pass moved into o3v_fused_policy_loss below (needs the rewards first)\n",
                          "\n"]
    return "".join(lines)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("path")
    ap.add_argument("--mode", choices=("calls", "fused"), default="calls")
    g = ap.add_mutually_exclusive_group()
    g.add_argument("--diff", action="store_true", help="print a unified diff (default)")
    g.add_argument("-o", "--output")
    g.add_argument("--in-place", action="store_true")
    a = ap.parse_args(argv)
    with open(a.path) as f:
        src = f.read()
    out = patch_source(src, a.mode)
    if a.output or a.in_place:
        with open(a.output or a.path, "w") as f:
            f.write(out)
    else:
        sys.stdout.writelines(difflib.unified_diff(src.splitlines(True), out.splitlines(True), a.path, a.path + " (o3v)"))
    return 0


if __name__ == "__main__":
    sys.exit(main())

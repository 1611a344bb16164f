"""Oracle: the CALL SEQUENCE of the reference's compute_loss around the hot path, and the fakes that let both the
real reference class and this restatement run without a checkpoint, a processor or video files.

Test infrastructure (see oracle/__init__.py).

`HostTrainer.compute_loss` restates what src/r1-v/src/open_r1/trainer/grpo_trainer.py does between the processor
call and the return (the part that involves the hot path), in the reference's order:
    :402-405  signature, ValueError on return_outputs
    :566-571  _prepare_inputs, prompt truncation
    :581-586  generate under unwrap_model_for_generation, prompt_length, completion_ids
    :590-596  EOS mask                                     (oracle/gspo.eos_mask)
    :598-609  pop ids / mask, repeat pixel_values
    :611-632  policy log-probs (with grad) then reference log-probs (inference_mode), each sliced
              `[:, prompt_length - 1:]`
    :635-636  KL, :639-656 decode + reward callables, :658-681 advantages, :691-706 objective,
    :711-738  metrics                                      (oracle/gspo.*)
Everything before :566 (chat template, video decode, key-frame interleave) is data plumbing outside the path and
is replaced by one processor call.  `tests/test_trainer_cpu.py` runs the REAL reference class (stub-imported,
oracle/ref_import.py) and this class on the same fakes and demands bit-equal loss and metrics, which is what
allows the GPU tests (the reference does not exist on the GPU box) to use this class as the host of the drop-in
mixin.  `HostTrainerPatched` is the same sequence after `integration/patch_reference.py` (level-2 integration).
"""
import contextlib
import types
from collections import defaultdict

import torch

from . import gspo as ogspo
from . import logps as ologps


# ----------------------------------------------------------------------------------------------- fakes
class FakeVLModel(torch.nn.Module):
    """A causal 'VL' model small enough for the CPU: embedding + causal running mean + MLP, `lm_head =
    nn.Linear(H, V, bias=False)` as the last layer (as in transformers' Qwen2_5_VLForConditionalGeneration), a
    `pixel_values` input that shifts the hidden states (so dropping the vision kwargs is visible), and a
    deterministic `generate`.  Parameters hold bf16-representable values and the final hidden states are rounded
    to bf16 values, so that a bf16 kernel and this fp32 model see identical numbers."""

    def __init__(self, vocab=1024, hidden=128, seed=0, eos_id=7, completion_len=24, num_generations=4):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        bf = lambda t: t.bfloat16().float()
        self.embed = torch.nn.Embedding(vocab, hidden)
        self.mix = torch.nn.Linear(hidden, hidden, bias=False)
        self.lm_head = torch.nn.Linear(hidden, vocab, bias=False)
        with torch.no_grad():
            self.embed.weight.copy_(bf(torch.randn(vocab, hidden, generator=g)))
            self.mix.weight.copy_(bf(torch.randn(hidden, hidden, generator=g) / hidden ** 0.5))
            self.lm_head.weight.copy_(bf(torch.randn(vocab, hidden, generator=g) * 0.05))
        self.vocab, self.eos_id, self.completion_len, self.G, self.seed = vocab, eos_id, completion_len, num_generations, seed
        self.warnings_issued = {}
        self.generate_calls = 0

    def backbone(self, input_ids, pixel_values=None):
        e = self.embed(input_ids)
        pos = torch.arange(1, e.shape[1] + 1, device=e.device, dtype=e.dtype).view(1, -1, 1)
        h = torch.tanh(self.mix(e + torch.cumsum(e, dim=1) / pos))
        if pixel_values is not None:
            h = h + 0.25 * torch.tanh(pixel_values.float().mean())
        return h.bfloat16().to(e.dtype)            # bf16-representable hidden states (differentiable casts)

    def forward(self, input_ids, attention_mask=None, pixel_values=None, image_grid_thw=None,
                pixel_values_videos=None, video_grid_thw=None, **kw):
        pv = pixel_values if pixel_values is not None else pixel_values_videos
        return types.SimpleNamespace(logits=self.lm_head(self.backbone(input_ids, pv)))

    @torch.no_grad()
    def generate(self, input_ids=None, attention_mask=None, generation_config=None, **kw):
        """[B, Lp] -> [B*G, Lp + Tc]: seeded 'samples' with one EOS planted per sequence (ids after it are
        arbitrary, as after a real generate with padding)."""
        self.generate_calls += 1
        B, G, Tc = input_ids.shape[0], self.G, self.completion_len
        g = torch.Generator().manual_seed(self.seed + 1000 + self.generate_calls)
        comp = torch.randint(8, self.vocab, (B * G, Tc), generator=g)
        lens = torch.randint(max(Tc // 3, 1), Tc + 1, (B * G,), generator=g)
        for n in range(B * G):
            if n % 5 != 4:                          # every fifth sequence never emits EOS
                comp[n, int(lens[n]) - 1] = self.eos_id
        return torch.cat([input_ids.repeat_interleave(G, dim=0), comp.to(input_ids.device)], dim=1)


class FakeProcessor:
    """Stands in for the Qwen2.5-VL processor: fixed prompt ids (left padded), a pixel tensor per image."""
    eos_token_id = 7
    pad_token_id = 0

    def __init__(self, prompt_len=40, vocab=1024, seed=3):
        self.prompt_len, self.vocab, self.seed = prompt_len, vocab, seed

    def __call__(self, text=None, images=None, videos=None, return_tensors="pt", padding=True, padding_side="left",
                 add_special_tokens=False, **kw):
        g = torch.Generator().manual_seed(self.seed + len(text[0]))
        B = len(text)
        ids = torch.randint(8, self.vocab, (B, self.prompt_len), generator=g)
        ids[:, :3] = self.pad_token_id
        mask = (ids != self.pad_token_id).long()
        out = dict(input_ids=ids, attention_mask=mask)
        if videos is not None:
            out["pixel_values_videos"] = torch.randn(16, 12, generator=g)
            out["video_grid_thw"] = torch.tensor([[1, 4, 4]])
            out["second_per_grid_ts"] = [0.5]
        else:
            out["pixel_values"] = torch.randn(16, 12, generator=g)
            out["image_grid_thw"] = torch.tensor([[1, 4, 4]])
        return out

    def batch_decode(self, ids, skip_special_tokens=True):
        return ["<think>%s</think><answer>%d</answer>" % (" ".join(str(int(t)) for t in row[:6]), int(row[0]))
                for row in ids]


class FakeAccelerator:
    def __init__(self, device):
        self.device = torch.device(device)

    def gather_for_metrics(self, t):
        return t

    def unwrap_model(self, model):
        while hasattr(model, "module") and isinstance(model.module, torch.nn.Module):
            model = model.module
        return model


def reward_len(prompts=None, completions=None, **kw):
    """Toy reward callables with the reference's signature (grpo_trainer.py:655)."""
    return [0.5 + 0.01 * (len(c[0]["content"]) % 7) for c in completions]


def reward_first(prompts=None, completions=None, **kw):
    return [float(int(c[0]["content"].split("<answer>")[1].split("<")[0]) % 3) for c in completions]


def make_example():
    """One dataset row as the reference's compute_loss expects it (`inputs` = [row], :407-416)."""
    return {"prompt": [{"role": "system", "content": "sys"},
                       {"role": "user", "content": [{"type": "image", "image": None}, {"type": "text", "text": "q?"}]}],
            "source": "gqa", "image_path": "x.jpg", "task": "visual QA", "answer": "[1,2,3,4]"}


def configure(trainer, model, ref_model, device="cpu", G=4, beta=0.04, gspo=True, max_prompt_length=32):
    """Fill the attributes compute_loss reads (:316-333, :569, :585 ...) on an instance made with __new__ (the
    real __init__ loads checkpoints)."""
    trainer.processing_class = FakeProcessor()
    trainer.accelerator = FakeAccelerator(device)
    trainer.args = types.SimpleNamespace(device=torch.device(device), past_index=-1, num_generations=G)
    trainer.is_deepspeed_enabled = False
    trainer.state = types.SimpleNamespace(global_step=2, max_steps=10)
    trainer.max_prompt_length = max_prompt_length
    trainer.num_generations = G
    trainer.generation_config = None
    trainer.ref_model = ref_model
    trainer.reward_funcs = [reward_len, reward_first]
    trainer.reward_processing_classes = [None, None]
    trainer.beta, trainer.epsilon_low, trainer.epsilon_high, trainer.gspo = beta, 0.2, 0.2, gspo
    trainer._metrics = defaultdict(list)
    trainer.model = model
    return trainer


def install_reference_fakes(module):
    """Give the stub-imported reference module (oracle/ref_import.load_trainer_class) working stand-ins for the
    trl / qwen_vl_utils helpers its compute_loss calls (:408, :453, :581, :640)."""
    from PIL import Image

    @contextlib.contextmanager
    def unwrap_model_for_generation(model, accelerator, **kw):
        yield accelerator.unwrap_model(model)

    module.maybe_apply_chat_template = lambda example, proc: {"prompt": "<|im_start|>user q?<|im_end|>"}
    module.is_conversational = lambda example: True
    module.unwrap_model_for_generation = unwrap_model_for_generation
    module.process_vision_info = lambda msgs, return_video_kwargs=False: ([Image.new("RGB", (64, 48))], None, {})


# ----------------------------------------------------------------------------------------------- restatement
class HostTrainer:
    """See the module docstring.  Duck-types the attributes `configure` sets."""

    def _get_per_token_logps(self, model, input_ids, **kwargs):              # :371-384
        logits = model(input_ids, **kwargs).logits
        logits = logits[:, :-1, :]
        input_ids = input_ids[:, 1:]
        per_token_logps = []
        for logits_row, input_ids_row in zip(logits, input_ids):
            log_probs = logits_row.log_softmax(dim=-1)
            per_token_logps.append(torch.gather(log_probs, dim=1, index=input_ids_row.unsqueeze(1)).squeeze(1))
        return torch.stack(per_token_logps)

    # -- pieces shared by the unpatched and the patched sequence
    def _rollout(self, model, inputs):
        device = self.accelerator.device
        prompts_text = ["<|im_start|>user q?<|im_end|>" for _ in inputs]
        inputs[0]["image_size_refine"] = (64, 48)                                               # :455
        inputs[0]["prompt_text_final"] = prompts_text[0]
        inputs[0]["step_percent"] = (self.state.global_step + 1) / self.state.max_steps          # :467-469
        prompt_inputs = self.processing_class(text=list(prompts_text), images=[None], videos=None, return_tensors="pt",
                                              padding=True, padding_side="left", add_special_tokens=False)
        prompt_inputs = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in prompt_inputs.items()}  # :566
        if self.max_prompt_length is not None:                                                   # :569-571
            prompt_inputs["input_ids"] = prompt_inputs["input_ids"][:, -self.max_prompt_length:]
            prompt_inputs["attention_mask"] = prompt_inputs["attention_mask"][:, -self.max_prompt_length:]
        prompt_ids = prompt_inputs["input_ids"]
        unwrapped = self.accelerator.unwrap_model(model)                                         # :581
        prompt_completion_ids = unwrapped.generate(**prompt_inputs, generation_config=self.generation_config)
        prompt_length = prompt_ids.size(1)                                                       # :583
        completion_ids = prompt_completion_ids[:, prompt_length:]                                # :585
        return prompt_inputs, prompt_completion_ids, prompt_length, completion_ids

    def _vision_kwargs(self, prompt_inputs, n):
        prompt_inputs.pop("input_ids")                                                           # :598-599
        prompt_inputs.pop("attention_mask")
        prompt_inputs["pixel_values"] = prompt_inputs["pixel_values"].repeat(n, 1)              # :601-603
        prompt_inputs["image_grid_thw"] = prompt_inputs["image_grid_thw"].repeat(n, 1)
        prompt_inputs.pop("second_per_grid_ts", None)                                            # :608-609
        return prompt_inputs

    def _ref_logps(self, model, prompt_completion_ids, prompt_length, prompt_inputs):
        with torch.inference_mode():                                                             # :619-626
            ref = self._get_per_token_logps(self.ref_model, prompt_completion_ids, **prompt_inputs)
            return ref[:, prompt_length - 1:]

    def _rewards(self, inputs, completion_ids, device):
        completions = self.processing_class.batch_decode(completion_ids, skip_special_tokens=True)     # :639-641
        completions = [[{"role": "assistant", "content": c}] for c in completions]
        prompts = [x["prompt"] for x in inputs for _ in range(self.num_generations)]             # :644
        rewards_per_func = torch.zeros(len(prompts), len(self.reward_funcs), device=device)      # :645
        for i, reward_func in enumerate(self.reward_funcs):                                      # :646-656
            kw = {k: [ex[k] for ex in inputs for _ in range(self.num_generations)]
                  for k in inputs[0].keys() if k not in ("prompt", "completion")}
            out = reward_func(prompts=prompts, completions=completions, **kw)
            rewards_per_func[:, i] = torch.tensor(out, dtype=torch.float32, device=device)
        return rewards_per_func

    def _metrics_block(self, completion_mask, rewards_per_func, rewards, std_grouped_rewards, mean_kl):
        g = self.accelerator.gather_for_metrics                                                  # :711-738
        self._metrics["completion_length"].append(g(completion_mask.sum(1)).float().mean().item())
        per_func = g(rewards_per_func).mean(0)
        for i, f in enumerate(self.reward_funcs):
            self._metrics["rewards/%s" % f.__name__].append(per_func[i].item())
        gathered = g(rewards)
        num_devices = gathered.size(0) // self.num_generations
        per_dev = gathered.view(num_devices, self.num_generations)
        self._metrics["all_wrong"].append((per_dev <= 1).all(dim=1).sum().item() / num_devices)
        self._metrics["all_correct"].append((per_dev >= 2).all(dim=1).sum().item() / num_devices)
        self._metrics["reward"].append(g(rewards).mean().item())
        self._metrics["reward_std"].append(g(std_grouped_rewards).mean().item())
        self._metrics["kl"].append(g(mean_kl).mean().item())

    def compute_loss(self, model, inputs, return_outputs=False, num_items_in_batch=None):
        if return_outputs:
            raise ValueError("The GRPOTrainer does not support returning outputs")               # :404-405
        device = self.accelerator.device
        prompt_inputs, pc_ids, prompt_length, completion_ids = self._rollout(model, inputs)
        _, completion_mask = ogspo.eos_mask(completion_ids.cpu(), self.processing_class.eos_token_id)   # :590-596
        completion_mask = completion_mask.to(device)
        prompt_inputs = self._vision_kwargs(prompt_inputs, len(pc_ids))
        per_token_logps = self._get_per_token_logps(model, pc_ids, **prompt_inputs)              # :612-613
        per_token_logps = per_token_logps[:, prompt_length - 1:]
        ref_per_token_logps = self._ref_logps(model, pc_ids, prompt_length, prompt_inputs)
        rewards_per_func = self._rewards(inputs, completion_ids, device)
        rewards, advantages, std = ogspo.group_advantages(rewards_per_func, self.num_generations)      # :658-681
        out = ogspo.gspo_objective(per_token_logps, ref_per_token_logps, completion_mask, advantages, self.beta,
                                   self.epsilon_low, self.epsilon_high, self.gspo)               # :635-636, 691-706
        self._metrics_block(completion_mask, rewards_per_func, rewards, std, out["mean_kl"])
        return out["loss"]


class HostTrainerPatched(HostTrainer):
    """The sequence after `integration/patch_reference.py --mode calls`: the EOS-mask block and the loss block are
    calls into the drop-in mixin (`o3v_completion_mask`, `compute_policy_loss`)."""

    def compute_loss(self, model, inputs, return_outputs=False, num_items_in_batch=None):
        if return_outputs:
            raise ValueError("The GRPOTrainer does not support returning outputs")
        device = self.accelerator.device
        prompt_inputs, pc_ids, prompt_length, completion_ids = self._rollout(model, inputs)
        completion_mask = self.o3v_completion_mask(completion_ids)
        prompt_inputs = self._vision_kwargs(prompt_inputs, len(pc_ids))
        per_token_logps = self._get_per_token_logps(model, pc_ids, **prompt_inputs)[:, prompt_length - 1:]
        ref_per_token_logps = self._ref_logps(model, pc_ids, prompt_length, prompt_inputs)
        rewards_per_func = self._rewards(inputs, completion_ids, device)
        return self.compute_policy_loss(per_token_logps, ref_per_token_logps, completion_mask, rewards_per_func)


class HostTrainerFused(HostTrainer):
    """The sequence after `integration/patch_reference.py --mode fused`: reference log-probs and rewards first,
    then ONE call that does the policy pass, the loss and their backward (`o3v_fused_policy_loss`)."""

    def compute_loss(self, model, inputs, return_outputs=False, num_items_in_batch=None):
        if return_outputs:
            raise ValueError("The GRPOTrainer does not support returning outputs")
        device = self.accelerator.device
        prompt_inputs, pc_ids, prompt_length, completion_ids = self._rollout(model, inputs)
        completion_mask = self.o3v_completion_mask(completion_ids)
        prompt_inputs = self._vision_kwargs(prompt_inputs, len(pc_ids))
        ref_per_token_logps = self._ref_logps(model, pc_ids, prompt_length, prompt_inputs)
        rewards_per_func = self._rewards(inputs, completion_ids, device)
        return self.o3v_fused_policy_loss(model, pc_ids, prompt_length, ref_per_token_logps, completion_mask,
                                          rewards_per_func, **prompt_inputs)

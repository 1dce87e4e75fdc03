"""(dev) chain path vs multi-kernel path, teacher-forced logits of the first steps, for one (size, B)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from whisper_trtllm_b200 import WhisperEngine, _abi, synthetic as synth

size, B, steps = sys.argv[1], int(sys.argv[2]), 4
cfg = synth.make_config(size, max_length=steps + 1)
sd = synth.make_weights(cfg, seed=17)
mel = synth.make_mel(B, seed=9).cuda()
out = {}
for chain in (0, 1):
    _abi.call("wb_set_decode_chain_path", chain)
    eng = WhisperEngine(cfg, sd, dtype="bfloat16", max_batch=B, enc_chunk=min(B, 16), device="cuda:0")
    if chain == 0:
        forced = eng.generate(mel).cpu().long()
    _, l = eng.generate(mel, forced_tokens=forced, dump_logits_steps=steps)
    out[chain] = l.float().cpu()
    eng.close()
errs = [float((out[1][s] - out[0][s]).abs().max() / out[0][s].abs().max()) for s in range(steps)]
rows = (out[1][0] - out[0][0]).abs().amax(dim=1) / out[0][0].abs().max()
print(f"{size} B={B} env BN={os.environ.get('WB_CHAIN_BN')} SPLITS={os.environ.get('WB_CHAIN_SPLITS')} SERIAL={os.environ.get('WB_CHAIN_SERIAL')}: "
      f"rel err per step {['%.4f' % e for e in errs]}; worst rows at step 0: {[(int(i), round(float(rows[i]), 3)) for i in rows.argsort(descending=True)[:4]]}")

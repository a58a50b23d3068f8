#!/bin/bash
# compute-sanitizer passes over one small training step (ResNet-18, 2 clips, TCN + language head + Adam) and the
# kernel-level stem / BatchNorm tests.  Usage on the GPU box: tools/sanitize.sh  -> gpurun_out/sanitize_*.log
set -u
mkdir -p gpurun_out
cat > /tmp/san_step.py <<'P'
import sys, torch
sys.path.insert(0, ".")
import r3m_b200
from r3m_b200 import R3M, Trainer
emb = torch.randn(2, 768)
r3m_b200.set_lang_encoder_factory(lambda dev: (lambda s: emb))
m = R3M("cuda", 1e-4, 1024, size=18, l2weight=1e-5, l1weight=1e-5, langweight=1.0, tcnweight=1.0)
model = torch.nn.DataParallel(m).cuda()
frames = torch.randint(0, 255, (2, 5, 3, 224, 224), device="cuda").float()
tr = Trainer(100)
for i in range(2):
    metrics, _ = tr.update(model, (frames, ["a", ""]), i)
print("step ok", metrics["full_loss"])
m.eval()
with torch.no_grad():
    for _ in range(3):
        out = model(frames[0])
print("eval ok", float(out.sum()))
P
for tool in ${TOOLS:-memcheck racecheck}; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 python /tmp/san_step.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|step ok|eval ok" gpurun_out/sanitize_$tool.log | tail -4
done

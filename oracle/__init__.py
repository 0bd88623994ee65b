"""CPU oracle for the corruption-sweep evaluation path  --  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the upstream project (Indra-jith/failure-aware-vision) ships no
implementation of this path: no classifier, no corruption
generators, no uncertainty / ECE / AUROC code and no tests that pin results.  This
package therefore *defines* the path as plain numpy / PyTorch-fp32 code following

  * README.md:22-24 of the reference (failure = wrong prediction with high confidence),
  * requirements.txt:1-6 of the reference (torch, torchvision, opencv, sklearn stack),
  * platform/backend/signal_analyzer.py:47-171 (frame layout BGR u8 HWC, return dict),
  * stock ``torchvision.models.resnet18/resnet50`` (topology + init),
  * the public definitions restated in SURVEY.md Appendix A (Philox4x32-10,
    Hendrycks & Dietterich corruptions, MC-dropout, ECE, bucketed AUROC).

What *is* pinned: Philox against the Random123 known-answer vectors, the ResNet
restatement against torchvision's own forward, AUROC against sklearn, the
SignalAnalyzer restatement against the reference's real code (golden vectors
generated from /root/reference by tests/golden/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  Product code (failure-aware-vision_b200/) never does.
"""

"""CPU oracle for the CXRMate SCST rollout path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``cxrmate_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker or the
timed CPU arm - never as the product path.

What it is: a plain-PyTorch (CPU, fp32) restatement of the reference's
algorithm for the hot path named in BASELINE.json:

* ``oracle.cvt``     - CvT-21 + projection head + multi-image regroup / mask
                       (reference modelling_longitudinal.py:28-90, HF
                       transformers 5.5.0 models/cvt/modeling_cvt.py:99-453).
* ``oracle.bert``    - 6-layer post-LN BERT decoder with cross-attention, LoRA on
                       self-attention Q/K, tied LM head, KV cache (HF
                       models/bert/modeling_bert.py:53-501; reference
                       modelling_longitudinal.py:163-170,173-249); and the
                       BERT-base encoder + CLS projection head of CXR-BERT.
* ``oracle.decode``  - the transformers-4.41-semantics decode loop the reference
                       was written for (reference
                       modelling_longitudinal.py:251-364; HF generation/utils.py
                       _sample, logits_process.py TopKLogitsWarper).
* ``oracle.reward``  - CXRBERTReward (reference tools/rewards/cxrbert.py:23-73).
* ``oracle.scst``    - sample()/reinforce_loss()/scst_step() (reference
                       modules/lightning_modules/longitudinal/scst/gen_prompt.py
                       :174-366).
* ``oracle.text``    - synthetic tokenizers and the section split
                       (reference modelling_longitudinal.py:413-513).
* ``oracle.weights`` - deterministic synthetic weights in the reference's
                       state_dict naming.

Parity pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against the reference's own
classes imported from /root/reference in the authoring container by
``oracle/pin_against_reference.py`` (which also writes ``tests/golden/*.npz``).
The CXR-BERT trunk is third-party hub code that is not vendored
(microsoft/BiomedVLP-CXR-BERT-specialized, no revision pin, unreachable
offline): that part is restated from its published architecture and is
"parity unpinned" against the hub implementation.
"""

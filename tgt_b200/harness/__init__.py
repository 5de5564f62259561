"""Bench / test harness around the hot path: PCQM-shaped synthetic batches and thin task models
(input embedding + heads) equivalent to the reference's lib/models/pcqm, which is *reused unchanged*
in a real deployment but does not exist on the GPU box.  Not part of the product path."""

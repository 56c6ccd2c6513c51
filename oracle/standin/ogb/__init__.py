"""Stand-in for the one piece of ``ogb`` (pinned ``ogb==1.2.2``, realworld_benchmark/environment_gpu.yml:43) the
reference's HIV / PCBA nets import: ``ogb.graphproppred.mol_encoder``.  TEST INFRASTRUCTURE (see oracle/__init__.py).
The package is absent and not installable here; the encoders are restated from the published release."""

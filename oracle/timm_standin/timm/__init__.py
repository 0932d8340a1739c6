"""Stand-in for timm==0.9.16 (not installed here; environment.yml:139 of the reference).
TEST INFRASTRUCTURE: lets /root/reference/train_options/models_original.py import unmodified
in the build container so golden vectors can be generated from the real reference."""
__version__ = "0.9.16-standin"

"""CPU oracle for the GSSD multibox hot path — TEST INFRASTRUCTURE, never imported by the product.

See gssd_oracle.c for the restated algorithm and oracle.py for the numpy-facing wrapper.
"""

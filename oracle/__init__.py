"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the SRDiff reverse-diffusion hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and there only as the checker / the timed CPU baseline.
The product package (``dgm_img_super_resolution_b200``) never imports it.
"""

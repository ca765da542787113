"""TEST INFRASTRUCTURE ONLY — CPU restatement of the VAENAR-TTS mel-synthesis hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and there only as the checker / the timed CPU baseline.
"""

"""TEST INFRASTRUCTURE.  The native reader raises ValueError("not a LAS file ...") where laspy raises LaspyException
(tests/integration_tests/test_io.py:152-158 expects an error of that name for a text file)."""
LaspyException = ValueError

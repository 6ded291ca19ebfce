"""Native layer of the B200 DRAM hot path: ctypes binding (`lib`), tensor wrappers (`ops`), autograd functions
(`functional`) and data-parallel plumbing (`dist`).  Importing this package does NOT load the shared library;
the first op does, and raises if it is missing (no fallback)."""

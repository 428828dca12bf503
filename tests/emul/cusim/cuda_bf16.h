// cusim: TEST INFRASTRUCTURE -- empty stand-in (nccl.h includes it; nothing of it is used)

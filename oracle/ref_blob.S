/* TEST INFRASTRUCTURE.  Embeds the reference's own compiled Fortran objects, byte for byte and straight from the
 * reference tree (REF_CU / REF_MODELS are absolute paths under /root/reference given by oracle/Makefile), into
 * oracle/_ref/libwmf_ref.so; see ref_loader.c. */
    .section .rodata
    .balign 64
    .globl wref_blob_cu
    .globl wref_blob_cu_end
wref_blob_cu:
    .incbin REF_CU
wref_blob_cu_end:
    .balign 64
    .globl wref_blob_models
    .globl wref_blob_models_end
wref_blob_models:
    .incbin REF_MODELS
wref_blob_models_end:
    .section .note.GNU-stack,"",@progbits

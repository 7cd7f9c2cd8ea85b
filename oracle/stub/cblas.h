/* empty stand-in for <cblas.h>: the reference includes it but the files built here never call it */

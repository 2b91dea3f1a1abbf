/* ORACLE shim: header included by the reference but no symbol used on the path */

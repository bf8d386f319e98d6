"""Self-tests of oracle/f77_to_c.py: the Fortran semantics the pin of the oracle relies on (expression typing, integer
division, integer powers, DO trip counts, array bounds and storage order, unformatted and formatted records), each on a
tiny fixed-form routine translated, compiled and called here."""
import ctypes as C
import importlib.util
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def f77():
    spec = importlib.util.spec_from_file_location("sos_f77_to_c", os.path.join(ROOT, "oracle", "f77_to_c.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def compile_fortran(f77, tmp_path, name, source):
    src = tmp_path / (name + ".F")
    src.write_text(source)
    c, protos, report = f77.translate_file(str(src), {})
    assert all(st == "ok" for _, st in report), report
    csrc = tmp_path / (name + ".c")
    csrc.write_text(f77.PRELUDE + "\n".join(protos) + "\n\n" + c)
    lib = tmp_path / ("lib" + name + ".so")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-w", "-shared", "-fPIC", "-o", str(lib), str(csrc), "-lm"], check=True)
    return C.CDLL(str(lib))


def test_expression_typing_and_powers(f77, tmp_path):
    lib = compile_fortran(f77, tmp_path, "typing", """
      SUBROUTINE TYPING(IS,L,A,OUT,KOUT)
      IMPLICIT NONE
      INTEGER*4 IS,L,KOUT(4)
      DOUBLE PRECISION A,OUT(10)
      REAL R
      OUT(1)=2.*IS/(L*(L+1.))
      OUT(2)=0.1
      OUT(3)=0.1D0
      OUT(4)=A**2
      OUT(5)=A**3
      OUT(6)=A**(-2)
      OUT(7)=A**(IS*0.5)
      R=1./3.
      OUT(8)=R
      OUT(9)=R*A
      OUT(10)=-A**2
      KOUT(1)=7/2
      KOUT(2)=(-7)/2
      KOUT(3)=IS**3
      KOUT(4)=2.9D0
      RETURN
      END
""")
    out, kout = np.zeros(10), np.zeros(4, dtype=np.int32)
    a = 1.2345678901234567
    lib.typing_(C.byref(C.c_int(3)), C.byref(C.c_int(7)), C.byref(C.c_double(a)), out.ctypes.data_as(C.POINTER(C.c_double)),
                kout.ctypes.data_as(C.POINTER(C.c_int)))
    f = np.float32
    assert out[0] == float(f(2.0) * f(3) / (f(7) * (f(7) + f(1.0))))             # all-REAL*4 expression, single-precision divide
    assert out[1] == float(f(0.1)) and out[2] == 0.1
    assert out[3] == a * a and out[4] == a * (a * a)                           # libgcc __powidf2 chains
    assert out[5] == 1.0 / (a * a)
    assert out[6] == a ** float(f(3) * f(0.5))                                 # REAL exponent, promoted operand by operand
    r = f(1.0) / f(3.0)
    assert out[7] == float(r) and out[8] == float(r) * a
    assert out[9] == -(a * a)                                                  # unary minus binds looser than **
    assert list(kout) == [3, -3, 27, 2]


def test_do_loops_arrays_and_records(f77, tmp_path):
    lib = compile_fortran(f77, tmp_path, "loops", """
      SUBROUTINE LOOPS(N,A,KOUT,FIC,IER)
      IMPLICIT NONE
      INTEGER*4 N,KOUT(6),IER,I,J,M,CNT
      DOUBLE PRECISION A(-2:2,0:3),B(-2:2,0:3),X(4)
      CHARACTER*200 FIC
      IER=0
      CNT=0
      M=N
      DO 10 I=1,M
         M=M+5
         CNT=CNT+1
 10   CONTINUE
      KOUT(1)=CNT
      KOUT(2)=I
      CNT=0
      DO I=10,1,-3
         CNT=CNT+1
      ENDDO
      KOUT(3)=CNT
      KOUT(4)=I
      CNT=0
      DO I=5,1
         CNT=CNT+1
      ENDDO
      KOUT(5)=CNT
      DO 20 J=0,3
      DO 20 I=-2,2
         IF (I.EQ.0) GOTO 20
         A(I,J)=10.D0*J+I
 20   CONTINUE
      OPEN(UNIT=11,FILE=FIC,FORM='UNFORMATTED',ERR=900)
      WRITE(11,err=900) ((A(I,J),I=-2,2),J=0,3),N
      WRITE(11,err=900) (A(I,1),I=2,-2,-1)
      CLOSE(11)
      OPEN(UNIT=11,FILE=FIC,FORM='UNFORMATTED',STATUS='OLD',ERR=900)
      READ(11,err=900) ((B(I,J),I=-2,2),J=0,3),M
      READ(11,err=900) (X(I),I=1,4)
      CLOSE(11)
      KOUT(6)=M
      DO J=0,3
         DO I=-2,2
            IF (B(I,J).NE.A(I,J)) IER=-1
         ENDDO
      ENDDO
      IF (X(1).NE.A(2,1)) IER=-2
      IF (X(4).NE.A(-1,1)) IER=-3
      RETURN
 900  IER=-9
      RETURN
      END
""")
    a = np.full(20, -1.0)
    kout = np.zeros(6, dtype=np.int32)
    ier = C.c_int(5)
    fic = str(tmp_path / "rec.bin")
    lib.loops_(C.byref(C.c_int(4)), a.ctypes.data_as(C.POINTER(C.c_double)), kout.ctypes.data_as(C.POINTER(C.c_int)),
               C.create_string_buffer(fic.encode().ljust(200), 200), C.byref(ier), C.c_size_t(200))
    assert ier.value == 0
    assert list(kout) == [4, 5, 4, -2, 0, 4]          # trip counts fixed at entry; DO variable after the loop; zero-trip loop
    ref = np.full((4, 5), -1.0)                       # [J][I+2]: column-major, first index fastest
    for j in range(4):
        for i in range(-2, 3):
            if i != 0:
                ref[j, i + 2] = 10.0 * j + i
    assert np.array_equal(a.reshape(4, 5), ref)
    raw = open(fic, "rb").read()
    assert np.frombuffer(raw, dtype=np.int32, count=1)[0] == 20 * 8 + 4        # gfortran record marker of the first record


def test_formatted_e15_8_round_trip(f77, pkg, tmp_path):
    lib = compile_fortran(f77, tmp_path, "fmt", """
      SUBROUTINE FMT(V,W,FIC,IER)
      IMPLICIT NONE
      DOUBLE PRECISION V(4),W(4)
      INTEGER*4 IER,K
      CHARACTER*200 FIC
      IER=0
      OPEN(UNIT=3,FILE=FIC,ERR=900)
      WRITE(3,207,err=900) V(1),V(2),V(3),V(4)
      CLOSE(3)
      OPEN(UNIT=3,FILE=FIC,STATUS='OLD',ERR=900)
      READ(3,207,err=900) W(1),W(2),W(3),W(4)
      CLOSE(3)
      RETURN
 900  IER=-1
      RETURN
 207  FORMAT(4(E15.8))
      END
""")
    v = np.array([0.123456789012, -3.14159265358979e-7, 0.0, 9.99999999e12])
    w = np.zeros(4)
    ier = C.c_int(9)
    fic = str(tmp_path / "fresnel.txt")
    lib.fmt_(v.ctypes.data_as(C.POINTER(C.c_double)), w.ctypes.data_as(C.POINTER(C.c_double)),
             C.create_string_buffer(fic.encode().ljust(200), 200), C.byref(ier), C.c_size_t(200))
    assert ier.value == 0
    fm = pkg.formats
    assert open(fic).read() == "".join(fm.fortran_e(x, 15, 8) for x in v) + "\n"
    assert np.array_equal(w, fm.round_e(v, 8))


def test_round_3_additions(f77, tmp_path):
    """What the aerosol routines needed: DINT / DNINT, dotted operators in lower case, a variable whose name starts with IF on the
    left of an assignment (not a logical IF), INDEX, formatted READ on a computed unit number, internal WRITE of an integer into a
    CHARACTER variable."""
    data = tmp_path / "tab.txt"
    data.write_text("   1.50000   2.25000\n   3.00000   4.00000\n")
    lib = compile_fortran(f77, tmp_path, "round3", """
      SUBROUTINE ROUND3(X,FIC,OUT,KOUT,CH)
      IMPLICIT NONE
      DOUBLE PRECISION X,OUT(8),A,B
      INTEGER*4 KOUT(4),IFIN,I,IGRANU
      CHARACTER*100 FIC
      CHARACTER*6 CH
      OUT(1)=DINT(X+X+20)
      OUT(2)=DNINT(X*1000.D+00)/1000.D+00
      OUT(3)=-DNINT(-X*10.D+00)/10.D+00
      OUT(4)=DNINT(2.5D+00)
      IGRANU=2
      KOUT(1)=0
      IF (IGRANU.eq.2) KOUT(1)=7
      IFIN=INDEX(FIC,' ')
      IFIN=IFIN-1
      KOUT(2)=IFIN
      IF (IFIN.le.0) KOUT(2)=-1
      OPEN(11,FILE=FIC,STATUS='OLD',ERR=99)
      OPEN(12,FILE=FIC,STATUS='OLD',ERR=99)
      DO I=1,2
         READ((I+10),555,ERR=99) A,B
         OUT(4+I)=A+B
      ENDDO
      READ(12,555,ERR=99) A,B
      OUT(7)=A*B
      CLOSE(11)
      CLOSE(12)
      WRITE(CH,'(I6)') INT(X*1000)
      DO I=1,6
         IF (CH(I:I).EQ.' ') CH(I:I)='0'
      ENDDO
      KOUT(3)=1
      RETURN
   99 KOUT(3)=-1
      RETURN
  555 FORMAT(2(1X,F9.5))
      END
""")
    out, kout = np.zeros(8), np.zeros(4, dtype=np.int32)
    ch = C.create_string_buffer(b"      ", 6)
    fic = C.create_string_buffer(str(data).encode().ljust(100), 100)
    lib.round3_(C.byref(C.c_double(1.4496)), fic, out.ctypes.data_as(C.POINTER(C.c_double)), kout.ctypes.data_as(C.POINTER(C.c_int)), ch,
                C.c_size_t(100), C.c_size_t(6))
    assert out[0] == 22.0 and out[1] == 1.45 and out[2] == 1.4 and out[3] == 3.0            # DNINT: halves away from zero
    assert kout[0] == 7 and kout[1] == len(str(data)) and kout[2] == 1
    assert out[4] == 3.75 and out[5] == 3.75                                                  # unit 11 line 1, unit 12 line 1
    assert out[6] == 12.0 and ch.raw == b"001449"                                             # unit 12 line 2; INT truncates 1449.6


def test_data_statements(f77, tmp_path):
    """DATA in the forms the reference uses (whole array, implied DO from 1, repeat counts, a scalar): REAL*4 constants stored in
    DOUBLE PRECISION elements keep their single-precision value, as DATA converts them."""
    lib = compile_fortran(f77, tmp_path, "datast", """
      SUBROUTINE DATAST(OUT,KOUT)
      PARAMETER(N=4)
      DOUBLE PRECISION OUT(12), A(N), B(3), S
      REAL R(2)
      INTEGER*4 KOUT(3), K(3)
      DATA A/
     C 1.013E+03, 0.1,
     C 2*2.5D0/
      DATA (B(I),I=1,3)/3.410E+22,18.,1.0E-06/
      DATA S/0.3/
      DATA R/0.7,1.5/
      DATA K/3*7/
      DO I=1,N
         OUT(I)=A(I)
      ENDDO
      DO I=1,3
         OUT(4+I)=B(I)
         KOUT(I)=K(I)
      ENDDO
      OUT(8)=S
      OUT(9)=R(1)
      OUT(10)=R(2)
      RETURN
      END
""")
    out, kout = np.zeros(12), np.zeros(3, dtype=np.int32)
    lib.datast_(out.ctypes.data_as(C.POINTER(C.c_double)), kout.ctypes.data_as(C.POINTER(C.c_int)))
    f = lambda x: float(np.float32(x))
    assert list(out[:10]) == [f(1.013e3), f(0.1), 2.5, 2.5, f(3.410e22), 18.0, f(1.0e-6), f(0.3), f(0.7), 1.5]
    assert list(kout) == [7, 7, 7]


def test_list_directed_read_and_a_format(f77, tmp_path):
    """List-directed numeric READ from a text file (values across records, D exponents, END=), a character constant as an actual
    argument, A and Hollerith fields on output, a dotted operator written with blanks."""
    lib = compile_fortran(f77, tmp_path, "listio", """
      SUBROUTINE LISTIO(FIN,FOUT,N,S)
      IMPLICIT NONE
      CHARACTER*100 FIN,FOUT
      INTEGER*4 N,K
      DOUBLE PRECISION S,V,W
      N=0
      S=0.D+00
      OPEN(1,FILE=FIN,STATUS='OLD',ERR=90)
   20 READ(1,*,ERR=90,END=30) V
      IF ( (V.GT.-1.D0).
     &      AND.(V.LT.1.D6) ) N=N+1
      S=S+V
      GOTO 20
   30 CLOSE(1)
      OPEN(1,FILE=FIN,STATUS='OLD',ERR=90)
      READ(1,*,ERR=90) K,V,W
      CLOSE(1)
      OPEN(2,FILE=FOUT,ERR=90)
      WRITE(2,100,ERR=90) N
      WRITE(2,120,ERR=90) FIN
      WRITE(2,130,ERR=90) K,W
      WRITE(2,140,ERR=90)
      CLOSE(2)
      CALL TAG("MIE",S)
      RETURN
   90 N=-1
      RETURN
  100 FORMAT(17hNB_TOTAL_ANGLES :,I4)
  120 FORMAT(6hFILE :,A)
  130 FORMAT(I3,1X,'W =',D12.5)
  140 FORMAT(12hINDEX  VALUE)
      END
      SUBROUTINE TAG(NAME,S)
      CHARACTER*3 NAME
      DOUBLE PRECISION S
      IF (NAME.EQ.'MIE') S=S+1000.D+00
      RETURN
      END
""")
    fin, fout = tmp_path / "in.txt", tmp_path / "out.txt"
    fin.write_text("3\n 2.5, 1.D+01\n\n4.25\n")
    n, s = C.c_int(0), C.c_double(0)
    pad = lambda p: C.create_string_buffer(str(p).encode().ljust(100), 100)
    lib.listio_(pad(fin), pad(fout), C.byref(n), C.byref(s), C.c_size_t(100), C.c_size_t(100))
    assert n.value == 3 and s.value == 3 + 2.5 + 4.25 + 1000.0         # one value per READ statement: the rest of a record is skipped
    out = fout.read_text().split("\n")
    assert out[0] == "NB_TOTAL_ANGLES :   3" and out[1].rstrip() == "FILE :" + str(fin) and len(out[1]) == 6 + 100
    assert out[2] == "  3 W = 0.10000D+02" and out[3] == "INDEX  VALUE"

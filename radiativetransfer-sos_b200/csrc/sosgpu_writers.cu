// sosgpu_writers.cu -- the ASCII result files of the drop-in surface: SOS_Up.txt / SOS_Down.txt as
// SOS_ABS_MAIN.F:2250-2519 writes them (record formats 55 / 56, :3095-3096) with the header blocks of SOS_OUTPUT_HEADER and
// SOS_OUTPUT_HEADER_POLAR_DIAG (SOS_TRPHI.F:1570-1796).  Host code: formatting text is not GPU work; the tables come from
// sosgpu_trphi_option / sosgpu_batch_trphi.  Byte compatibility is tested against header text derived from the reference's
// own WRITE statements (tests/make_golden_headers.py) and an independent formatter of the records.
#include "../../include/sosgpu.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

namespace {

// Fortran Aw output of a literal: leftmost w bytes, or right-justified when shorter (gfortran counts BYTES: the UTF-8 degree
// sign of the reference's source counts two, which truncates two header lines -- reproduced as is)
std::string fa(const char *s, size_t w)
{
  const size_t n = strlen(s);
  if (n >= w) return std::string(s, w);
  return std::string(w - n, ' ') + s;
}
std::string ff(double x, int w, int d)
{
  char b[64];
  snprintf(b, sizeof b, "%*.*f", w, d, x);
  if ((int)strlen(b) > w) return std::string(w, '*');
  return b;
}
// Ew.d: 0.dddddE+ee, correctly rounded (glibc printf), right-justified
std::string fe(double x, int w, int d)
{
  char m[64], s[80];
  if (x == 0.0) snprintf(s, sizeof s, "0.%0*dE+00", d, 0);
  else {
    snprintf(m, sizeof m, "%.*e", d - 1, std::fabs(x));
    char *e = strchr(m, 'e');
    const int ex = atoi(e + 1) + 1;
    *e = 0;
    std::string dig;
    for (const char *p = m; *p; ++p) if (*p != '.') dig += *p;
    if (std::abs(ex) < 100) snprintf(s, sizeof s, "%s0.%sE%c%02d", x < 0 ? "-" : "", dig.c_str(), ex >= 0 ? '+' : '-', std::abs(ex));
    else snprintf(s, sizeof s, "%s0.%s%c%03d", x < 0 ? "-" : "", dig.c_str(), ex >= 0 ? '+' : '-', std::abs(ex));
  }
  std::string r = s;
  if ((int)r.size() > w && r.compare(0, 2, "0.") == 0) r = r.substr(1);            // gfortran drops the optional zero when tight
  else if ((int)r.size() > w && r.compare(0, 3, "-0.") == 0) r = "-" + r.substr(2);
  if ((int)r.size() > w) return std::string(w, '*');
  return std::string(w - r.size(), ' ') + r;
}

const char *RULE = "#-----------------------------------------------------------------------------------------------------";

void columns_block(FILE *f)
{
  fprintf(f, "%s\n", fa("#   SCA_ANG :  Scattering angle (in degrees)", 44).c_str());
  fprintf(f, "%s\n", fa("#   I       :  Stokes parameter I at output altitude z (in sr-1)", 64).c_str());
  fprintf(f, "%s\n", fa("#              normalised to the extraterrestrial solar irradiance (PI * L(z) / Esun)", 85).c_str());
  fprintf(f, "%s\n", fa("#   Q       :  Stokes parameter Q at output altitude z (in sr-1)", 64).c_str());
  fprintf(f, "%s\n", fa("#              normalised to the extraterrestrial solar irradiance", 66).c_str());
  fprintf(f, "%s\n", fa("#   U       :  Stokes parameter U at output altitude z (in sr-1)", 64).c_str());
  fprintf(f, "%s\n", fa("#              normalised to the extraterrestrial solar irradiance ", 66).c_str());
  fprintf(f, "%s\n", fa("#   POL_ANG :  Polarization angle (in degrees).  Note: if undefined the value is -999.00", 88).c_str());
  fprintf(f, "%s\n", fa("#   POL_RATE:  Degree of polarization (in %)", 44).c_str());
  fprintf(f, "%s\n", fa("#   IPOL    :  Polarized intensity at level z (in sr-1)", 55).c_str());
  fprintf(f, "%s\n", fa("#              normalised to the extraterrestrial solar irradiance (PI * Lpol(z) / Esun)", 88).c_str());
  fprintf(f, "%s\n", fa(RULE, 102).c_str());
}

// SOS_OUTPUT_HEADER (SOS_TRPHI.F:1570-1700)
void header_view1(FILE *f, int updown, double phi1, double phi2, double zalt)
{
  if (updown == 1) fprintf(f, "%s\n", fa("# UPWARD RADIANCE FIELD VERSUS THE VIEWING ZENITH ANGLE", 55).c_str());
  else fprintf(f, "%s\n", fa("# DOWNWARD RADIANCE FIELD VERSUS THE VIEWING ZENITH ANGLE", 57).c_str());
  fprintf(f, "%s\n", fa("# (RELATIVE AZIMUTH AND ALTITUDE ARE FIXED)", 43).c_str());
  fprintf(f, "%s\n", fa(RULE, 102).c_str());
  fprintf(f, "%s\n", fa("# Relative azimuth (degrees) :", 30).c_str());
  fprintf(f, "%s\n", fa("#", 1).c_str());
  fprintf(f, "%s\n", fa("#      Relative azimuth convention :", 36).c_str());
  if (updown == 1) {
    fprintf(f, "%s\n", fa("#        180\xc2\xb0 <-> Satellite and Sun in the same half-plane", 58).c_str());
    fprintf(f, "%s\n", fa("#          0\xc2\xb0 <-> Satellite and Sun in opposite half-planes with respect to the zenith direction", 96).c_str());
  } else {
    fprintf(f, "%s\n", fa("#        180\xc2\xb0 <-> Viewing direction and Sun in the same half-plane", 66).c_str());
    fprintf(f, "%s\n", fa("#          0\xc2\xb0 <-> Viewing direction and Sun in opposite half-planes with respect to the zenith direction", 104).c_str());
  }
  fprintf(f, "%s\n", fa("#", 1).c_str());
  fprintf(f, "%s\n", fa("#      Simulated relative azimuth (degrees) :", 45).c_str());
  fprintf(f, "%s %s\n", fa("#          for VZA < 0 (sign convention):", 41).c_str(), ff(phi2, 7, 3).c_str());
  fprintf(f, "%s %s\n", fa("#          for VZA > 0 (sign convention):", 41).c_str(), ff(phi1, 7, 3).c_str());
  fprintf(f, "%s\n", fa("#", 1).c_str());
  fprintf(f, "%s %s\n", fa("# Value of the selected altitude for the output (km) :", 54).c_str(), ff(zalt, 7, 3).c_str());
  fprintf(f, "%s\n", fa("#", 1).c_str());
  fprintf(f, "%s\n", fa("# Columns parameters :", 22).c_str());
  fprintf(f, "%s\n", fa("#   VZA     :  Viewing Zenith Angle (in degrees)", 48).c_str());
  columns_block(f);
  fprintf(f, "%s\n", fa("#   VZA     SCA_ANG       I              Q              U         POL_ANG  POL_RATE    IPOL", 91).c_str());
  fprintf(f, "%s\n", fa("#(degrees) (degrees)    (sr-1)         (sr-1)         (sr-1)      (degrees)  (%)      (sr-1)", 92).c_str());
}

// SOS_OUTPUT_HEADER_POLAR_DIAG (SOS_TRPHI.F:1705-1796)
void header_view2(FILE *f, int updown, double zalt)
{
  if (updown == 1) fprintf(f, "%s\n", fa("#UPWARD RADIANCE FIELD VERSUS THE AZIMUTH ANGLE AND VIEWING ZENITH ANGLE", 72).c_str());
  else fprintf(f, "%s\n", fa("#DOWNWARD RADIANCE FIELD VERSUS THE AZIMUTH ANGLE AND VIEWING ZENITH ANGLE", 74).c_str());
  fprintf(f, "%s\n", fa("#(ALTITUDE FIXED)", 17).c_str());
  fprintf(f, "%s\n", fa(RULE, 102).c_str());
  fprintf(f, "%s\n", fa("# Relative azimuth convention :", 31).c_str());
  fprintf(f, "%s\n", fa("#        180\xc2\xb0 <-> Satellite and Sun in the same half-plane", 58).c_str());
  fprintf(f, "%s\n", fa("#          0\xc2\xb0 <-> Satellite and Sun in opposite half-planes with respect to the zenith direction", 96).c_str());
  fprintf(f, "%s\n", fa("#", 1).c_str());
  fprintf(f, "%s %s\n", fa("# Value of the selected altitude for the output (km) :", 54).c_str(), ff(zalt, 7, 3).c_str());
  fprintf(f, "%s\n", fa("#", 1).c_str());
  fprintf(f, "%s\n", fa("# Columns parameters :", 22).c_str());
  fprintf(f, "%s\n", fa("#   PHI     :  Relative azimuth Angle (in degrees)", 50).c_str());
  fprintf(f, "%s\n", fa("#   VZA     :  Viewing Zenith Angle (in degrees)", 48).c_str());
  columns_block(f);
  fprintf(f, "%s\n", fa("#   PHI      VZA     SCA_ANG        I              Q              U       POL_ANG  POL_RATE    IPOL", 99).c_str());
  fprintf(f, "%s\n", fa("#(degrees) (degrees) (degrees)    (sr-1)         (sr-1)         (sr-1)    (degrees)  (%)      (sr-1)", 100).c_str());
}

// FORMAT 55: 2(2X,F7.2),2X,3(E13.6,2X),2(F7.2,2X),E13.6
void rec55(FILE *f, double vza, double sca, double xi, double xq, double xu, double xan, double tpol, double lpol)
{
  fprintf(f, "  %s  %s  %s  %s  %s  %s  %s  %s\n", ff(vza, 7, 2).c_str(), ff(sca, 7, 2).c_str(), fe(xi, 13, 6).c_str(), fe(xq, 13, 6).c_str(),
          fe(xu, 13, 6).c_str(), ff(xan, 7, 2).c_str(), ff(tpol, 7, 2).c_str(), fe(lpol, 13, 6).c_str());
}
// FORMAT 56: 3(2X,F7.2),1X,3(E13.6,2X),2(1X,F7.2),E13.6
void rec56(FILE *f, double phi, double vza, double sca, double xi, double xq, double xu, double xan, double tpol, double lpol)
{
  fprintf(f, "  %s  %s  %s %s  %s  %s   %s %s%s\n", ff(phi, 7, 2).c_str(), ff(vza, 7, 2).c_str(), ff(sca, 7, 2).c_str(), fe(xi, 13, 6).c_str(),
          fe(xq, 13, 6).c_str(), fe(xu, 13, 6).c_str(), ff(xan, 7, 2).c_str(), ff(tpol, 7, 2).c_str(), fe(lpol, 13, 6).c_str());
}

}  // namespace

// Tables as sosgpu_trphi_option returns them: [7][nphi_cap][N] in the order SCA, I, Q, U, POL_ANG, POL_RATE, L_POL.
extern "C" int sosgpu_write_updown(const char *fic_up, const char *fic_down, int nbmu, int itrphi, double phios, int pas_phi, double zout,
                                   const double *phi_fin, const double *theta_fin, const double *up, const double *down, int nphi_cap,
                                   int fix_sca_index)
{
  if (!fic_up || !fic_down || !theta_fin || !up || !down || nbmu < 1 || (itrphi != 1 && itrphi != 2)) return SOSGPU_ERR_ARG;
  const int N = nbmu;
  const int nphi = itrphi == 1 ? 2 : (pas_phi >= 1 ? 360 / pas_phi + 1 : 0);
  if (nphi < 1 || nphi > nphi_cap || (itrphi == 2 && !phi_fin)) return SOSGPU_ERR_ARG;
  FILE *fu = fopen(fic_up, "w"), *fd = fopen(fic_down, "w");
  if (!fu || !fd) { if (fu) fclose(fu); if (fd) fclose(fd); return SOSGPU_ERR_IER; }
  auto T = [&](const double *tab, int t, int ip, int jj) { return tab[((size_t)t * nphi_cap + ip) * N + jj]; };
  const double alt_up = (zout == -1) ? 120.0 : zout, alt_dn = (zout == -1) ? 0.0 : zout;   // CTE_TOA_ALT = 120 km (SOS.h:197)
  if (itrphi == 1) {                                             // SOS_ABS_MAIN.F:2262-2390
    header_view1(fu, 1, phios, phios + 180.0, alt_up);
    header_view1(fd, 2, phios, phios + 180.0, alt_dn);
    for (int j = N; j >= 1; --j) {                               // half-plane PHIOS + 180: negative viewing angles
      const int jj = j - 1;
      rec55(fu, -theta_fin[jj], T(up, 0, 0, jj), T(up, 1, 0, jj), T(up, 2, 0, jj), T(up, 3, 0, jj), T(up, 4, 0, jj), T(up, 5, 0, jj), T(up, 6, 0, jj));
      rec55(fd, -theta_fin[jj], T(down, 0, 0, jj), T(down, 1, 0, jj), T(down, 2, 0, jj), T(down, 3, 0, jj), T(down, 4, 0, jj), T(down, 5, 0, jj), T(down, 6, 0, jj));
    }
    for (int j = 1; j <= N; ++j) {
      const int jj = j - 1;
      rec55(fu, theta_fin[jj], T(up, 0, 1, jj), T(up, 1, 1, jj), T(up, 2, 1, jj), T(up, 3, 1, jj), T(up, 4, 1, jj), T(up, 5, 1, jj), T(up, 6, 1, jj));
      rec55(fd, theta_fin[jj], T(down, 0, 1, jj), T(down, 1, 1, jj), T(down, 2, 1, jj), T(down, 3, 1, jj), T(down, 4, 1, jj), T(down, 5, 1, jj), T(down, 6, 1, jj));
    }
  } else {                                                       // SOS_ABS_MAIN.F:2395-2470
    header_view2(fu, 1, alt_up);
    header_view2(fd, 2, alt_dn);
    int ip = 0;
    for (int iphi = 0; iphi <= 360; iphi += pas_phi, ++ip)
      for (int jj = 0; jj < N; ++jj) {
        // The reference indexes the upward scattering angle by the azimuth in DEGREES (SCA_UP_FIN(IPHI,JJ), :2467) instead of
        // the azimuth counter IP; entries it never filled are zero.  Reproduced unless fix_sca_index is set.
        const double sca_up = fix_sca_index ? T(up, 0, ip, jj) : (iphi < nphi ? T(up, 0, iphi, jj) : 0.0);
        rec56(fu, phi_fin[ip], theta_fin[jj], sca_up, T(up, 1, ip, jj), T(up, 2, ip, jj), T(up, 3, ip, jj), T(up, 4, ip, jj), T(up, 5, ip, jj), T(up, 6, ip, jj));
        rec56(fd, phi_fin[ip], theta_fin[jj], T(down, 0, ip, jj), T(down, 1, ip, jj), T(down, 2, ip, jj), T(down, 3, ip, jj), T(down, 4, ip, jj), T(down, 5, ip, jj), T(down, 6, ip, jj));
      }
  }
  const bool ok = !ferror(fu) && !ferror(fd);
  fclose(fu); fclose(fd);
  return ok ? SOSGPU_OK : SOSGPU_ERR_IER;
}

// ------------------------------------------------------------------------------------------------
// The two optional ASCII files SOS_PROC writes after the synthesis (SOS_PROC.F:3779-3874): transmissions (-SOS.Trans) and
// fluxes (-SOS.Flux).  The formatted records follow FORMAT 1005 / 1006 / 1010 / 2010 / 2016-2018 / 2020 (:4944-4951).  The
// list-directed records (WRITE(u,*)) are laid out as gfortran lays them out -- one leading blank, character items as they are,
// a REAL*8 item as a separator blank + G25.17E3 (F form between 0.1 and 1e17: right-justified in 20 columns + 5 blanks, else
// 1PE25.17E3) -- from the documented rules, not from a run: there is no Fortran compiler here to pin those few lines against.
namespace {
std::string list_real8(double x)
{
  char b[80];
  const double ax = std::fabs(x);
  if (x == 0.0) return std::string("   0.0000000000000000     ");
  if (ax >= 0.1 && ax < 1e17) {                                    // F editing: 17 significant digits
    snprintf(b, sizeof b, "%.16e", ax);
    const int k = atoi(strchr(b, 'e') + 1) + 1;                     // 10^(k-1) <= ax < 10^k after rounding to 17 digits
    snprintf(b, sizeof b, "%s%.*f", x < 0 ? "-" : "", 17 - k, ax);
    std::string f(b);
    if (f.size() < 20) f = std::string(20 - f.size(), ' ') + f;
    return " " + f + "     ";
  }
  snprintf(b, sizeof b, "%.16E", ax);                              // d.ddddddddddddddddE+ee -> three exponent digits
  char *e = strchr(b, 'E');
  const int ex = atoi(e + 1);
  *e = 0;
  char out[96];
  snprintf(out, sizeof out, "%s%sE%c%03d", x < 0 ? "-" : "", b, ex < 0 ? '-' : '+', abs(ex));
  std::string f(out);
  if (f.size() < 25) f = std::string(25 - f.size(), ' ') + f;
  return " " + f;
}
}  // namespace

extern "C" int sosgpu_write_trans(const char *fictrans, double tetas, double ttot_tronc, double ttot_vrai, double tdifmus, int nbmu,
                                  const double *rmu, const double *tdifmug)
{
  if (!fictrans || !rmu || !tdifmug || nbmu < 1) return SOSGPU_ERR_ARG;
  FILE *f = fopen(fictrans, "w");
  if (!f) return SOSGPU_ERR_IER;
  const double pi = std::acos(-1.0);
  double tdir_tronc = std::exp(-ttot_tronc / std::cos(pi * tetas / 180.0));
  double tdir_vrai = std::exp(-ttot_vrai / std::cos(pi * tetas / 180.0));
  fprintf(f, "Solar Zenith Angle : %s\n", ff(tetas, 7, 3).c_str());
  fprintf(f, "Direct transmission TOA -> surface : %s\n", ff(tdir_vrai, 8, 4).c_str());
  fprintf(f, "  \n");
  fprintf(f, " Diffuse transmittance : TOA -> surface\n");
  fprintf(f, "    thetas = %s   td(thetas) = %s\n", ff(tetas, 6, 3).c_str(), ff(tdifmus + tdir_tronc - tdir_vrai, 7, 4).c_str());
  fprintf(f, "  \n");
  fprintf(f, " Diffuse transmittance : surface -> TOA\n");
  for (int j = 1; j <= nbmu; ++j) {                                // rmu[j-1] = RMU(J), tdifmug[j-1] = TDIFMUG(J)
    tdir_tronc = std::exp(-ttot_tronc / rmu[j - 1]);
    tdir_vrai = std::exp(-ttot_vrai / rmu[j - 1]);
    fprintf(f, "    thetav = %s   td(thetav) = %s\n", ff(std::acos(rmu[j - 1]) * 180.0 / pi, 6, 3).c_str(),
            ff(tdifmug[j - 1] + tdir_tronc - tdir_vrai, 7, 4).c_str());
  }
  return fclose(f) == 0 ? SOSGPU_OK : SOSGPU_ERR_IER;
}

extern "C" int sosgpu_write_flux(const char *ficflux, double tetas, double ttot_tronc, double ttot_vrai, double emoins, double eplus,
                                 double tr, double hr, double ta, double ha, const double *zalt, const double *tauabs,
                                 double *tdir_vrai_out, double *flux_diff_down_out, double *flux_down_out)
{
  const double pi = std::acos(-1.0);
  const double tdir_tronc = std::exp(-ttot_tronc / std::cos(pi * tetas / 180.0));
  const double tdir_vrai = std::exp(-ttot_vrai / std::cos(pi * tetas / 180.0));
  const double flux_diff_down = emoins + tdir_tronc - tdir_vrai, flux_down = emoins + tdir_tronc;
  if (tdir_vrai_out) *tdir_vrai_out = tdir_vrai;                   // SOS_PROC returns them whether or not the file is written
  if (flux_diff_down_out) *flux_diff_down_out = flux_diff_down;
  if (flux_down_out) *flux_down_out = flux_down;
  if (!ficflux || strcmp(ficflux, "NO_OUTPUT") == 0) return SOSGPU_OK;
  if (!zalt || !tauabs) return SOSGPU_ERR_ARG;
  FILE *f = fopen(ficflux, "w");
  if (!f) return SOSGPU_ERR_IER;
  fprintf(f, "Solar Zenith Angle : %s\n", ff(tetas, 7, 3).c_str());
  fprintf(f, "  \n");
  fprintf(f, " Downward fluxes at BOA (normalized by TOA solar flux)\n");
  fprintf(f, "   - Downward direct flux at BOA : %s\n", ff(tdir_vrai, 9, 5).c_str());
  fprintf(f, "   - Downward diffuse flux at BOA: %s\n", ff(flux_diff_down, 9, 5).c_str());
  fprintf(f, "   ==> Downward total flux at BOA: %s\n", ff(flux_down, 9, 5).c_str());
  fprintf(f, "  \n");
  fprintf(f, " Upward diffuse flux at TOA (normalized by TOA solar flux):%s\n", list_real8(eplus).c_str());
  fprintf(f, "\n\n");
  fprintf(f, " According to the following profile\n");
  fprintf(f, " Z(km)    MOT     AOT     GOT     TOTAL\n");
  for (int i = 50; i >= 1; --i) {                                  // zalt[i-1] = ABS_USERPROFIL(I,1); tauabs[50-i] = TAUABS(51-I)
    const double z = zalt[i - 1];
    const double tr_z = tr * std::exp(-z / hr), ta_z = ta * std::exp(-z / ha), tg_z = tauabs[50 - i];
    fprintf(f, "%s  %s %s %s %s\n", ff(z, 7, 2).c_str(), ff(tr_z, 7, 4).c_str(), ff(ta_z, 7, 4).c_str(), ff(tg_z, 7, 4).c_str(),
            ff(tr_z + ta_z + tg_z, 7, 4).c_str());
  }
  return fclose(f) == 0 ? SOSGPU_OK : SOSGPU_ERR_IER;
}
